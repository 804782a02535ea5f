// ingest.cu -- row (f)3 of SURVEY.md section 8: record split of raw, uncompressed FASTQ / FASTA text on the device.
//
// Replaces the parsing half of ReadParserParallel::parse (src/Commons.hpp:5846-5911: one thread reads records with
// kseq inside an `omp critical` and hands a copy of every Read to a worker): the text block crosses PCIe as it is,
// a newline index is built on the device, every record's sequence (and quality) line is located from it, and the
// sequence bytes go straight from the text into the 2-bit layout of the packed sketch kernel -- no host-side
// per-character work at all.  Supported: 4-line FASTQ and 2-line FASTA (sequence on one line, what PacBio / ONT
// base callers write), '\n' or '\r\n' line ends.  Anything else is reported so that the caller falls back to kseq.
#include "common.cuh"
#include "engine.cuh"

namespace mdbg {

constexpr int NL_THREADS = 256, NL_BYTES = 16, NL_TILE = NL_THREADS * NL_BYTES;

__device__ __forceinline__ uint32_t newline_mask16(const uint8_t* text, uint64_t n, uint64_t pos) {
    // bit j = 1 when text[pos + j] == '\n' (bytes at / after n do not count)
    uint32_t m = 0;
    if (pos + 16 <= n && ((uintptr_t)(text + pos) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + pos);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t x = w[q] ^ 0x0A0A0A0Au;                                  // zero byte <=> newline
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; // 0x80 in every zero byte
            m |= ((z * 0x00204081u) >> 28) << (4 * q);
        }
    } else {
        for (int j = 0; j < 16; j++)
            if (pos + j < n && text[pos + j] == '\n') m |= 1u << j;
    }
    return m;
}

__global__ void __launch_bounds__(NL_THREADS) newline_count_kernel(const uint8_t* text, uint64_t n, uint32_t* counts) {
    const uint64_t pos = (uint64_t)blockIdx.x * NL_TILE + (uint64_t)threadIdx.x * NL_BYTES;
    uint32_t c = pos < n ? __popc(newline_mask16(text, n, pos)) : 0u;
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __shared__ uint32_t ws[NL_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < NL_THREADS / 32; i++) t += ws[i];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(NL_THREADS) newline_write_kernel(const uint8_t* text, uint64_t n, const uint64_t* tile_off,
                                                                   uint64_t* nl_pos) {
    const uint64_t pos = (uint64_t)blockIdx.x * NL_TILE + (uint64_t)threadIdx.x * NL_BYTES;
    const uint32_t m = pos < n ? newline_mask16(text, n, pos) : 0u;
    const uint32_t c = __popc(m);
    const uint32_t incl = warp_inclusive_scan(c);
    __shared__ uint32_t ws[NL_THREADS / 32];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = incl - c;
    for (uint32_t i = 0; i < (threadIdx.x >> 5); i++) before += ws[i];
    uint64_t dst = tile_off[blockIdx.x] + before;
    uint32_t rest = m;
    while (rest) {
        const uint32_t j = __ffs(rest) - 1;
        rest &= rest - 1;
        nl_pos[dst++] = pos + j;
    }
}

// One thread per record.  Line t of the text is (nl[t-1] + 1 .. nl[t]) with nl[-1] = -1; a record is `lines` lines.
__global__ void __launch_bounds__(256) fastx_records_kernel(const FastxArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_records) return;
    const uint64_t t0 = i * a.lines;
    const uint64_t h_lo = t0 ? a.nl[t0 - 1] + 1 : 0;                    // header line
    const uint64_t s_lo = a.nl[t0] + 1;
    uint64_t s_hi = a.nl[t0 + 1];                                        // sequence line [s_lo, s_hi)
    bool bad = a.text[h_lo] != (a.lines == 4 ? '@' : '>');
    if (s_hi > s_lo && a.text[s_hi - 1] == '\r') s_hi--;
    uint64_t q_lo = 0;
    if (a.lines == 4) {
        const uint64_t p_lo = a.nl[t0 + 1] + 1;
        q_lo = a.nl[t0 + 2] + 1;
        uint64_t q_hi = a.nl[t0 + 3];
        if (q_hi > q_lo && a.text[q_hi - 1] == '\r') q_hi--;
        bad |= a.text[p_lo] != '+' || (q_hi - q_lo) != (s_hi - s_lo);
    }
    bad |= (s_hi - s_lo) >> 31 != 0;
    a.seq_start[i] = s_lo;
    a.seq_len[i] = bad ? 0u : (uint32_t)(s_hi - s_lo);
    if (a.qual_start) a.qual_start[i] = q_lo;
    if (bad) atomicAdd(a.n_bad, 1ULL);
}

void launch_newline_count(const uint8_t* text, uint64_t n, uint32_t* counts, cudaStream_t s) {
    if (n == 0) return;
    newline_count_kernel<<<(unsigned)((n + NL_TILE - 1) / NL_TILE), NL_THREADS, 0, s>>>(text, n, counts);
}
void launch_newline_write(const uint8_t* text, uint64_t n, const uint64_t* tile_off, uint64_t* nl_pos, cudaStream_t s) {
    if (n == 0) return;
    newline_write_kernel<<<(unsigned)((n + NL_TILE - 1) / NL_TILE), NL_THREADS, 0, s>>>(text, n, tile_off, nl_pos);
}
uint64_t newline_tiles(uint64_t n) { return (n + NL_TILE - 1) / NL_TILE; }
void launch_fastx_records(const FastxArgs& a, cudaStream_t s) {
    if (a.n_records == 0) return;
    fastx_records_kernel<<<(unsigned)((a.n_records + 255) / 256), 256, 0, s>>>(a);
}

}  // namespace mdbg
