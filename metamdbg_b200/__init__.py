"""metamdbg_b200 -- Blackwell-native minimizer-sketch + k-min-mer count engine
behind a C ABI (include/mdbg_b200.h).  The product is libmdbg_b200.so; this
package is its thin host-side mirror of the reference interface."""
from .engine import CountTable, Engine, KminmerCounter, MdbgError, MinimizerParser, Sketch  # noqa: F401
from .multik import multi_k_sweep  # noqa: F401,E402
