// CPU unit test of the host 2-bit packer (metamdbg_b200/csrc/pack_host.cpp): every clean read round-trips through
// the packed form with code (c >> 1) & 3, every read holding a byte outside "ACGT" is spilled verbatim.
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../metamdbg_b200/csrc/pack_host.hpp"

using namespace mdbg;

int main() {
    std::mt19937 rng(7);
    const char* al = "ACGT";
    std::vector<std::string> reads;
    for (int i = 0; i < 3000; i++) {
        const int n = (i < 200) ? i : (int)(rng() % 20000);       // every short length incl. 0, 15, 16, 17, 31, 32, 33, 63..65
        std::string s;
        for (int j = 0; j < n; j++) s += al[rng() & 3];
        if (i % 97 == 0 && n > 5) s[n / 2] = 'N';
        if (i % 131 == 0 && n > 5) s[1] = 'a';
        if (i % 211 == 0 && n > 70) s[n - 1] = '#';
        reads.push_back(s);
    }
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offs{0}, pk{0};
    for (auto& s : reads) {
        bases.insert(bases.end(), s.begin(), s.end());
        offs.push_back(bases.size());
        pk.push_back(pk.back() + (s.size() + 15) / 16);
    }
    std::vector<uint32_t> pack(pk.back() + 16, 0xFFFFFFFFu);
    std::vector<uint64_t> src(reads.size());
    std::vector<uint8_t> asc(bases.size() + 16 * reads.size() + 64);
    std::atomic<uint64_t> cur{0};
    HostPool* pool = host_pool_create(5);
    host_pack_reads(pool, bases.data(), offs.data(), 0, (uint32_t)reads.size() / 2, pk.data(), pack.data(), src.data(),
                    asc.data(), &cur);                               // two ranges, like two pipeline pieces
    host_pack_reads(pool, bases.data(), offs.data(), (uint32_t)reads.size() / 2, (uint32_t)reads.size(), pk.data(),
                    pack.data(), src.data(), asc.data(), &cur);
    // the two-halves form the host-batch path uses: piece i+1 is already being packed while piece i is consumed
    std::vector<uint32_t> pack2(pk.back() + 16, 0xFFFFFFFFu);
    std::vector<uint64_t> src2(reads.size());
    std::vector<uint8_t> asc2(asc.size());
    std::atomic<uint64_t> cur2{0};
    int bad_overlap = 0;
    {
        const uint32_t n_pieces = 16, per = (uint32_t)reads.size() / n_pieces;
        auto lo = [&](uint32_t p) { return p * per; };
        auto hi = [&](uint32_t p) { return p + 1 == n_pieces ? (uint32_t)reads.size() : (p + 1) * per; };
        PackJob* job = host_pack_start(pool, bases.data(), offs.data(), lo(0), hi(0), pk.data(), pack2.data(), src2.data(),
                                       asc2.data(), &cur2);
        for (uint32_t p = 0; p < n_pieces; p++) {
            host_pack_wait(pool, job);
            job = nullptr;
            const uint64_t asc_end = cur2.load();
            if (p + 1 < n_pieces)
                job = host_pack_start(pool, bases.data(), offs.data(), lo(p + 1), hi(p + 1), pk.data(), pack2.data(),
                                      src2.data(), asc2.data(), &cur2);
            for (uint32_t r = lo(p); r < hi(p); r++) {                // "enqueue piece p": its data must be final by now
                if (src2[r] >> 63) { if ((src2[r] & ~(1ull << 63)) + reads[r].size() > asc_end) bad_overlap++; continue; }
                for (size_t w = 0; w < (reads[r].size() + 15) / 16; w++) {
                    uint32_t want = 0;
                    for (size_t j = 16 * w; j < reads[r].size() && j < 16 * w + 16; j++)
                        want |= (((unsigned char)reads[r][j] >> 1) & 3u) << (2 * (j % 16));
                    if (pack2[pk[r] + w] != want) { bad_overlap++; break; }
                }
            }
        }
    }
    host_pool_destroy(pool);
    int bad = bad_overlap, dirty = 0;
    for (size_t r = 0; r < reads.size(); r++) {
        const std::string& s = reads[r];
        bool clean = true;
        for (char c : s) if (c != 'A' && c != 'C' && c != 'G' && c != 'T') clean = false;
        if (src[r] >> 63) {
            dirty++;
            if (clean) bad++;
            if (memcmp(asc.data() + (src[r] & ~(1ull << 63)), s.data(), s.size()) != 0) bad++;
            continue;
        }
        if (!clean || src[r] != pk[r]) { bad++; continue; }
        // exact words: base j at bits [2j, 2j+1] of word j/16, padding bits of the last word zero
        for (size_t w = 0; w < (s.size() + 15) / 16; w++) {
            uint32_t want = 0;
            for (size_t j = 16 * w; j < s.size() && j < 16 * w + 16; j++)
                want |= (((unsigned char)s[j] >> 1) & 3u) << (2 * (j % 16));
            if (pack[pk[r] + w] != want) { bad++; break; }
        }
    }
    for (size_t w = pk.back(); w < pack.size(); w++) if (pack[w] != 0xFFFFFFFFu) bad++;    // nothing written past the end
    printf("reads %zu dirty %d bad %d threads_default %d isa %s\n", reads.size(), dirty, bad, host_default_threads(), host_pack_isa());
    return bad == 0 && dirty > 20 ? 0 : 1;
}
