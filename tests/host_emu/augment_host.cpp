// TEST INFRASTRUCTURE: the phases of the augmentation kernel (csrc/augment_core.cuh) compiled for the HOST with g++
// (-ffp-contract=off) and run serially, one band after the other - the same source the CUDA kernel is built from, so the
// CPU suite checks its arithmetic against the oracle without a GPU.  Never loaded by the product.
#include <cstring>
#include <vector>

#include "augment_core.cuh"

using namespace kp::aug;

static const FilterDef kFilters[6] = KP_AUG_FILTER_TABLE;

extern "C" int kp_emu_augment_frames(const unsigned char* src, const kp_frame_plan* plans, int n_frames, float* out) {
    std::vector<uint8_t> tile(ROWS * ROWB), res(BAND * ROWB);
    float lut[256], kf[25];
    for (int v = 0; v < 256; ++v) lut[v] = model_range(v);
    for (int f = 0; f < n_frames; ++f) {
        const kp_frame_plan& plan = plans[f];
        for (int band = 0; band < S / BAND; ++band) {
            float* o = out + (static_cast<long long>(f) * S + band * BAND) * ROWB;
            if (plan.zero) {
                for (int i = 0; i < BAND * ROWB; ++i) o[i] = lut[0];
                continue;
            }
            const int fid = plan.filter_id;
            if (fid >= 0 && fid <= 6)
                for (int i = 0; i < 25; ++i) kf[i] = filter_tap(kFilters, fid, i);
            unsigned int lsum = 0;
            // three "threads" of different strides exercise the (tid, nthr) indexing
            for (int t = 0; t < 3; ++t) phase_gather(src, plan, band, tile.data(), t, 3);
            if (fid == 9)
                for (int t = 0; t < 3; ++t) lsum += phase_luma(src, plan, t, 3);
            for (int t = 0; t < 3; ++t) phase_filter(tile.data(), plan, kf, lsum, band, res.data(), t, 3);
            for (int i = 0; i < BAND * ROWB; ++i) o[i] = lut[res[i]];
        }
    }
    return 0;
}
