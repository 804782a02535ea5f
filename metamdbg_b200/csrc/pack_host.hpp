// pack_host.hpp -- host-side 2-bit packing of ASCII reads before they cross PCIe (transfer compression only:
// the packed bases are unpacked again inside the sketch kernel; no part of the result is computed here).
#pragma once

#include <atomic>
#include <cstdint>

namespace mdbg {

class HostPool;                                   // persistent worker threads
int host_default_threads();                       // CPUs this process can really use (cgroup quota, ranks per node)
HostPool* host_pool_create(int n_threads);        // n_threads <= 0: MDBG_HOST_THREADS or host_default_threads()
int host_pool_size(const HostPool* p);
void host_pool_destroy(HostPool* p);
const char* host_pack_isa();                      // "avx512" | "avx2" | "scalar": the packer the CPU (and MDBG_PACK_ISA) selects
// one read on the calling thread: ceil(len / 16) words; false (words undefined) when it holds a byte outside "ACGT"
bool host_pack_one(const uint8_t* bases, uint64_t len, uint32_t* words_out);

// Pack reads [r0, r1) of an ASCII batch.  Read r goes to pack_out[pk_off[r] ...] (16 bases per u32, base j at bits
// [2j, 2j+1], code (c >> 1) & 3) and src_out[r] = pk_off[r]; a read holding any byte outside "ACGT" is instead
// copied verbatim to asc_out (16-byte aligned slot reserved from *asc_cursor) and src_out[r] = (1 << 63) | offset.
void host_pack_reads(HostPool* pool, const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1,
                     const uint64_t* pk_off, uint32_t* pack_out, uint64_t* src_out, uint8_t* asc_out,
                     std::atomic<uint64_t>* asc_cursor);

// The same in two halves: host_pack_start hands the piece to the workers and returns, host_pack_wait blocks until
// it is packed (and frees the job).  One job per pool at a time.  The host-batch path uses this to enqueue piece
// i's copies and kernel launches (~20 driver calls) while piece i+1 is already being packed.
struct PackJob;
PackJob* host_pack_start(HostPool* pool, const uint8_t* bases, const uint64_t* offsets, uint32_t r0, uint32_t r1,
                         const uint64_t* pk_off, uint32_t* pack_out, uint64_t* src_out, uint8_t* asc_out,
                         std::atomic<uint64_t>* asc_cursor);
void host_pack_wait(HostPool* pool, PackJob* job);

}  // namespace mdbg
