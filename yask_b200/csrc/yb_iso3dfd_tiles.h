// Table entry describing one compiled variant of the iso3dfd tiled kernels (yb_iso3dfd.cu / yb_iso3dfd_radii.cu).
#pragma once
#include "yb_iso3dfd.cuh"

namespace yb {

typedef void (*IsoKernelFn)(const IsoMaps, const IsoParams);
struct TileCfg {
    bool fused_ok;      // kernel can store boundary planes into the x neighbours (gen2 PW/U variants)
    const char* name;
    int ty, tz, hp, hrows, threads;
    uint32_t smem;
    IsoKernelFn fn[4];  // per FP mode (3 = debug memory-only probe, gen2 only)
};


// The default variant (tile 16x128, producer warpgroup, 2 planes per trip) instantiated for radius 1..7;
// radius 8 has the full variant table in yb_iso3dfd.cu.  Returns nullptr for other radii.
const TileCfg* iso_radius_cfg(int radius);

}  // namespace yb
