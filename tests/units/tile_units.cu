// Host-side unit test of device helpers that are plain integer logic (compiled by nvcc, run on the CPU; no CUDA
// API call is made).  Built and run by tests/test_units_cpu.py with -DSPHB_FLAT_STAGING=1.
//
// tile_entry_of_slot (tile.cuh, flat staging): for random segment tables -- empty segments anywhere, as k_tile_setup
// produces them for grid edges and non-seam tiles -- the slot -> entry map must be exactly the enumeration of
// SPHB_TILE_FOR_STAGED (slot = seg.off + (e - seg.b) for e in [seg.b, seg.e), segments in order).
#include <cstdio>
#include <cstdlib>

#include "../../sph-poiseuille-flow_b200/csrc/tile.cuh"

using namespace sphb;

int main()
{
    unsigned long long rng = 0x9e3779b97f4a7c15ull;
    auto next = [&]() {
        rng ^= rng << 13;
        rng ^= rng >> 7;
        rng ^= rng << 17;
        return rng;
    };
    long long checked = 0;
    for (int trial = 0; trial < 20000; ++trial) {
        TileDesc T;
        int off = 0, cursor = (int)(next() % 1000);
        for (int k = 0; k < kStageRows * 3; ++k) {
            const unsigned r = (unsigned)(next() % 100);
            // ~half of the segments empty (wrap columns of interior tiles, rows outside the grid)
            const int len = r < 50 ? 0 : (int)(next() % (r < 60 ? 3 : 130));
            cursor += (int)(next() % 50);            // entry ranges of different rows are apart
            T.seg[k].b = cursor;
            T.seg[k].e = cursor + len;
            T.seg[k].off = off;
            off += len;
            cursor += len;
        }
        T.n_staged = off;
        for (int k = 0; k < kStageRows * 3; ++k)
            for (int e = T.seg[k].b; e < T.seg[k].e; ++e) {
                const int slot = T.seg[k].off + (e - T.seg[k].b);
                const int got = tile_entry_of_slot(T, slot);
                if (got != e) {
                    std::printf("FAIL trial %d segment %d slot %d: entry %d, expected %d\n", trial, k, slot, got, e);
                    return 1;
                }
                ++checked;
            }
    }
    std::printf("OK %lld slots\n", checked);
    return 0;
}
