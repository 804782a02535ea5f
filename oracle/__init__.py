"""TEST-ONLY CPU oracle for the metaMDBG sketch+count path (see mdbg_oracle.h)."""
