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
    host_pool_destroy(pool);
    int bad = 0, dirty = 0;
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
