// Host-side check of the weight-table / split-row index maps (tilingnn_b200/csrc/layouts.cuh, hsplit.cuh): every map
// must be a bijection onto its table, otherwise two weights would share a slot.  Built and run by tests/test_layouts_host.py.
#include <cstdio>
#include <set>
#include <vector>

#include "hsplit.cuh"
#include "layouts.cuh"

using namespace tgnn;

static int fails = 0;
static void expect_perm(const std::vector<size_t>& idx, size_t size, const char* what) {
    std::set<size_t> s(idx.begin(), idx.end());
    bool ok = s.size() == idx.size() && idx.size() == size && *s.rbegin() == size - 1;
    printf("%-44s %s (%zu entries)\n", what, ok ? "ok" : "FAIL", idx.size());
    if (!ok) ++fails;
}

int main() {
    const int kmaps[3] = {KMAP_GATHER, KMAP_NATURAL, KMAP_CHAIN}, nmaps[2] = {NMAP_NATURAL, NMAP_CONTIG8};
    const int shapes[3][2] = {{32, 32}, {32, 64}, {64, 32}};
    for (auto& sh : shapes)
        for (int km : kmaps)
            for (int nm : nmaps) {
                if (nm == NMAP_CONTIG8 && sh[1] != 32) continue;
                if (km == KMAP_GATHER && sh[0] != 32) continue;
                std::vector<size_t> v;
                for (int k = 0; k < sh[0]; ++k)
                    for (int n = 0; n < sh[1]; ++n)
                        for (int hl = 0; hl < 2; ++hl) v.push_back(frag_index(k, n, sh[1], km, nm, hl));
                char name[64];
                snprintf(name, sizeof name, "frag_index %dx%d kmap %d nmap %d", sh[0], sh[1], km, nm);
                expect_perm(v, (size_t)2 * sh[0] * sh[1], name);
            }
    {
        std::vector<size_t> v;
        for (int k = 0; k < 32; ++k)
            for (int n = 0; n < 32; ++n)
                for (int hl = 0; hl < 2; ++hl) v.push_back((size_t)hfrag_half_index(k, n, hl));
        expect_perm(v, 2048, "hfrag_half_index 32x32");
    }
    for (auto& sh : shapes)
        for (int nm : nmaps) {
            if (nm == NMAP_CONTIG8 && sh[1] != 32) continue;
            std::vector<size_t> v;
            for (int k = 0; k < sh[0]; ++k)
                for (int n = 0; n < sh[1]; ++n)
                    for (int hl = 0; hl < 2; ++hl) v.push_back(hfrag_nat_half_index(k, n, sh[1], nm, hl));
            char name[64];
            snprintf(name, sizeof name, "hfrag_nat_half_index %dx%d nmap %d", sh[0], sh[1], nm);
            expect_perm(v, (size_t)2 * sh[0] * sh[1], name);
        }
    {
        std::vector<size_t> v;
        for (int n = 0; n < 32; ++n)
            for (int k = 0; k < 32; ++k) v.push_back((size_t)tile_pos(n, k));
        expect_perm(v, 1024, "tile_pos 32x32 (SWIZZLE_128B image)");
    }
    {
        std::vector<size_t> v;
        for (int q = 0; q < 8; ++q) v.push_back((size_t)xh_pos(q));
        expect_perm(v, 8, "xh_pos");
        // lane t's two pieces (q = t, q = 4 + t) must be adjacent and 32-byte aligned
        for (int t = 0; t < 4; ++t)
            if (xh_pos(t) != 2 * t || xh_pos(4 + t) != 2 * t + 1) { printf("xh_pos adjacency FAIL at t=%d\n", t); ++fails; }
    }
    return fails ? 1 : 0;
}
