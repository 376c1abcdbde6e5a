// iso3dfd engine: spec (what the reference compiler derives from
// /root/reference/src/stencils/Iso3dfdStencil.cpp), FD coefficients,
// TMA tensor maps and kernel launch logic.
#include <math.h>

#include <algorithm>
#include <vector>
#include <stdio.h>
#include <stdlib.h>

#include "yb_core.h"
#include "yb_iso3dfd.cuh"
#include "yb_iso3dfd_tiles.h"

namespace yb {

// ---- FD coefficients ----------------------------------------------------------------------------------
// Fornberg's recurrence with the same operation order as the reference
// (/root/reference/src/contrib/coefficients/fd_coeff.cpp:53-101, called through
// get_center_fd_coefficients, /root/reference/src/common/fd_coeff2.cpp:49-57), then the x3 centre
// weight and /50^2 scaling of Iso3dfdStencil.cpp:68-88, then the 16-significant-digit print/parse
// round trip every constant takes through the reference's generated source (Cpp.cpp:39-53).
static void center_fd_coeffs(int order, int radius, std::vector<double>& out) {
    const int np = 2 * radius + 1;
    std::vector<double> pts(np);
    for (int i = 0; i < np; i++) pts[i] = double(i - radius);
    // w[m][n][k]: weight of point k for derivative m using points 0..n
    std::vector<double> w(size_t(order + 1) * np * np, 0.0);
    auto W = [&](int m, int n, int k) -> double& { return w[(size_t(m) * np + n) * np + k]; };
    W(0, 0, 0) = 1.0;
    double c1 = 1.0;
    for (int n = 1; n < np; n++) {
        double c2 = 1.0;
        const int mmax = std::min(n, order);
        for (int k = 0; k < n; k++) {
            const double c3 = pts[n] - pts[k];
            c2 = c2 * c3;
            for (int m = 0; m <= mmax; m++) {
                double t = (pts[n] - 0.0) * W(m, n - 1, k);
                if (m > 0) t -= m * W(m - 1, n - 1, k);
                t *= 1.0 / c3;
                if (t == 0.0) t = 0.0;
                W(m, n, k) = t;
            }
        }
        for (int m = 0; m <= mmax; m++) {
            double t = 0.0;
            if (m > 0) t += m * W(m - 1, n - 1, n - 1);
            t -= (pts[n - 1] - 0.0) * W(m, n - 1, n - 1);
            t *= c1 / c2;
            if (t == 0.0) t = 0.0;
            W(m, n, n) = t;
        }
        c1 = c2;
    }
    out.resize(np);
    for (int i = 0; i < np; i++) out[i] = W(order, np - 1, i);
}

static double const_roundtrip(double v) {
    if (double(int(v)) == v) return v;
    char buf[64];
    snprintf(buf, sizeof buf, "%.15e", v);
    return strtod(buf, nullptr);
}

void iso3dfd_coeffs(int radius, double* c) {
    std::vector<double> full;
    center_fd_coeffs(2, radius, full);
    const double d2 = 50.0 * 50.0;
    for (int i = 0; i <= 2 * radius; i++) {
        if (i == radius) full[i] *= 3.0;
        full[i] /= d2;
    }
    for (int r = 0; r <= radius; r++) c[r] = const_roundtrip(full[radius + r]);
}

StencilSpec iso3dfd_spec(int radius, int elem_bytes, bool) {
    StencilSpec sp;
    sp.name = "iso3dfd";
    sp.description = "isotropic 3-D finite-difference wave equation, 2nd order in time, order 2*radius in space";
    sp.domain_dims = {"x", "y", "z"};
    sp.radius = radius;
    sp.elem_bytes = elem_bytes;
    VarSpec p;
    p.name = "p";
    p.step_alloc = 2;   // write-back of t+1 over t-1
    p.is_output = true;
    p.l1_norm = 1;
    DimSpec t; t.name = "t"; t.kind = DIM_STEP;
    p.dims.push_back(t);
    const char* dn[3] = {"x", "y", "z"};
    for (int d = 0; d < 3; d++) {
        DimSpec ds; ds.name = dn[d]; ds.kind = DIM_DOMAIN; ds.domain_index = d; ds.halo_l = ds.halo_r = radius;
        p.dims.push_back(ds);
    }
    VarSpec v;
    v.name = "v";
    v.step_alloc = 1;
    for (int d = 0; d < 3; d++) {
        DimSpec ds; ds.name = dn[d]; ds.kind = DIM_DOMAIN; ds.domain_index = d;
        v.dims.push_back(ds);
    }
    sp.vars = {p, v};
    StageSpec st;
    st.name = "stage_1";
    st.outputs = {0};
    st.inputs = {0, 1};
    // reference compiler counts for radius r: reads 6r+3, writes 1, fp ops 7r+5 (r=8: 51/1/61)
    st.reads = 6 * radius + 3;
    st.writes = 1;
    st.fp_ops = 7 * radius + 5;
    sp.stages = {st};
    return sp;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D tiled map over one step slot of a var (dims x,y,z; z unit stride) with box (bz, by, 1).
int make_map(CUtensorMap* map, const Var& v, int slot, int bz, int by) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return set_error(YB_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    const Dim* dx = v.domain_dim(0);
    const Dim* dy = v.domain_dim(1);
    const Dim* dz = v.domain_dim(2);
    cuuint64_t gdim[3] = {cuuint64_t(dz->alloc), cuuint64_t(dy->alloc), cuuint64_t(dx->alloc)};
    cuuint64_t gstr[2] = {cuuint64_t(dy->stride) * 4, cuuint64_t(dx->stride) * 4};
    cuuint32_t box[3] = {cuuint32_t(bz), cuuint32_t(by), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, v.slot_ptr(slot), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(YB_ECUDA, "cuTensorMapEncodeTiled failed with code %d", int(r));
    return 0;
}

}  // namespace

// 3-D tiled map over one step slot of a full-rank var of any element size (generated sweep kernels, yb_gen.cu).
int make_var_tensor_map(void* map_, const Var& v, int slot, int box_z, int box_y) {
    CUtensorMap* map = static_cast<CUtensorMap*>(map_);
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return set_error(YB_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    const Dim* dx = v.domain_dim(0);
    const Dim* dy = v.domain_dim(1);
    const Dim* dz = v.domain_dim(2);
    if (!dx || !dy || !dz || dz->stride != 1) return set_error(YB_EUNSUPPORTED, "tensor map: var '%s' is not a full-rank 3-D var", v.spec.name.c_str());
    const cuuint64_t eb = cuuint64_t(v.elem_bytes);
    cuuint64_t gdim[3] = {cuuint64_t(dz->alloc), cuuint64_t(dy->alloc), cuuint64_t(dx->alloc)};
    cuuint64_t gstr[2] = {cuuint64_t(dy->stride) * eb, cuuint64_t(dx->stride) * eb};
    cuuint32_t box[3] = {cuuint32_t(box_z), cuuint32_t(box_y), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, v.elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, v.slot_ptr(slot), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(YB_ECUDA, "cuTensorMapEncodeTiled failed with code %d (var '%s')", int(r), v.spec.name.c_str());
    return 0;
}

namespace {

template <class T, int PW, int U>
TileCfg cfg_tile(const char* name) {
    return TileCfg{true, name, T::TY, T::TZ, T::HP, T::HROWS, T::THREADS + 128 * PW, T::SMEM_BYTES,
                   {iso3dfd_tma2_kernel<T, 0, PW, U>, iso3dfd_tma2_kernel<T, 1, PW, U>, iso3dfd_tma2_kernel<T, 2, PW, U>,
                    iso3dfd_tma2_kernel<T, 3, PW, U>}};
}

// Compiled variants of the sweep kernel for radius 8 (option `tile`): tile shape x producer warpgroup x planes per trip.
constexpr int NTILES = 8;
constexpr int DEFAULT_TILE = 7;
const TileCfg& tile_cfg(int i) {
    static const TileCfg cfgs[NTILES] = {
        cfg_tile<IsoTile2<8, 16, 16, 5>, 0, 1>("32x64, in-loop producer"),
        cfg_tile<IsoTile2<8, 8, 32, 5>, 0, 1>("16x128, in-loop producer"),
        cfg_tile<IsoTile2<8, 16, 16, 5>, 1, 1>("32x64 + producer warpgroup"),
        cfg_tile<IsoTile2<8, 16, 16, 5>, 0, 2>("32x64, 2 planes per trip"),
        cfg_tile<IsoTile2<8, 8, 32, 5>, 1, 1>("16x128 + producer warpgroup"),
        cfg_tile<IsoTile2<8, 8, 32, 5>, 0, 2>("16x128, 2 planes per trip"),
        cfg_tile<IsoTile2<8, 16, 16, 5>, 1, 2>("32x64 + producer warpgroup, 2 planes per trip"),
        cfg_tile<IsoTile2<8, 8, 32, 5>, 1, 2>("16x128 + producer warpgroup, 2 planes per trip"),
    };
    return cfgs[i];
}

struct IsoEngine : Engine {
    int radius = 8;
    double coef[ISO_MAX_R + 1] = {0};
    std::string kernel = "auto";   // auto | tma | direct
    int tile = DEFAULT_TILE;       // index into tile_cfg(): 16x128, producer warpgroup, 2 planes per trip
    int lx = 0;                    // planes per sweep chunk (0 = choose per launch)
    int grid_override = 0;
    int num_sms = 148;
    bool attr_set[NTILES][4] = {};
    // L2 policy of the sweep (profiles/r2_iso3dfd.md): both p(t) streams evict_last (a plane is fetched halo-less 8 sweep steps
    // before it is fetched with its halo, and its halo rows are the neighbouring tiles' centre rows), p(t-1) / v evict_first
    // (streamed once), results leave with streaming stores: -0.4 GB of DRAM reads per 1024^3 launch, +0.7 % (same-box A/B)
    int pol_c = 2, pol_h = 2, pol_pv = 1, st_cs = 1;
    bool mem_probe = false;      // debug: fp_mode=3 style memory-only kernel
    int peer_probe = 0;          // debug: see IsoParams::peer_probe
    IsoMaps maps[NTILES][2];       // [tile][cur slot]
    bool maps_ok = false;

    int set_option(Solution&, const std::string& k, const std::string& v) override {
        if (k == "kernel") {
            if (v != "auto" && v != "tma" && v != "direct") return YB_EINVAL;
            kernel = v;
        } else if (k == "tile") {
            int t = atoi(v.c_str());
            if (t < 0 || t >= NTILES) return YB_EINVAL;
            tile = t;
        }
        else if (k == "lx") { lx = std::max(0, atoi(v.c_str())); }
        else if (k == "grid") { grid_override = std::max(0, atoi(v.c_str())); }
        else if (k == "mem_probe") { mem_probe = atoi(v.c_str()) != 0; }
        else if (k == "peer_probe") { peer_probe = atoi(v.c_str()); }
        else if (k == "pol_c") { pol_c = atoi(v.c_str()); }
        else if (k == "pol_h") { pol_h = atoi(v.c_str()); }
        else if (k == "pol_pv") { pol_pv = atoi(v.c_str()); }
        else if (k == "st_cs") { st_cs = atoi(v.c_str()) != 0; }
        else return YB_EINVAL;
        return 0;
    }
    bool get_option(const Solution&, const std::string& k, std::string& v) const override {
        if (k == "kernel") v = kernel;
        else if (k == "tile") v = std::to_string(tile);
        else if (k == "lx") v = std::to_string(lx);
        else if (k == "grid") v = std::to_string(grid_override);
        else return false;
        return true;
    }

    int prepare(Solution& s) override {
        radius = s.spec.radius;
        if (radius < 1 || radius > ISO_MAX_R) return set_error(YB_EUNSUPPORTED, "iso3dfd radius must be in 1..%d", ISO_MAX_R);
        if (s.spec.elem_bytes != 4) return set_error(YB_EUNSUPPORTED, "iso3dfd: only 4-byte elements are implemented in this round");
        iso3dfd_coeffs(radius, coef);
        cudaDeviceProp prop;
        YB_CUDA(cudaGetDeviceProperties(&prop, s.device));
        num_sms = prop.multiProcessorCount;
        maps_ok = false;
        if (prop.major >= 9 && (radius == 8 || iso_radius_cfg(radius))) {
            const Var& p = s.vars[0];
            const Var& v = s.vars[1];
            for (int tl = 0; tl < (radius == 8 ? NTILES : 1); tl++) {
                const TileCfg& c = radius == 8 ? tile_cfg(tl) : *iso_radius_cfg(radius);
                for (int cur = 0; cur < 2; cur++) {
                    IsoMaps& m = maps[tl][cur];
                    if (int rc = make_map(&m.h, p, cur, c.hp, c.hrows)) return rc;
                    if (int rc = make_map(&m.c, p, cur, c.tz, c.ty)) return rc;
                    if (int rc = make_map(&m.p, p, 1 - cur, c.tz, c.ty)) return rc;
                    if (int rc = make_map(&m.v, v, 0, c.tz, c.ty)) return rc;
                }
            }
            maps_ok = true;
            const TileCfg& c = radius == 8 ? tile_cfg(tile) : *iso_radius_cfg(radius);
            for (int m = 0; m < 4; m++) preload_kernel((const void*)c.fn[m]);
        }
        preload_kernel((const void*)iso3dfd_direct_kernel<0>);
        preload_kernel((const void*)iso3dfd_direct_kernel<1>);
        preload_kernel((const void*)iso3dfd_direct_kernel<2>);
        return 0;
    }

    void fill_params(const Solution& s, int cur, const Box& box, IsoParams& P) const {
        const Var& p = s.vars[0];
        const Var& v = s.vars[1];
        const Dim *px = p.domain_dim(0), *py = p.domain_dim(1), *pz = p.domain_dim(2);
        const Dim *vx = v.domain_dim(0), *vy = v.domain_dim(1), *vz = v.domain_dim(2);
        P.out = reinterpret_cast<float*>(p.slot_ptr(1 - cur)) + p.origin_offset();
        P.cur = reinterpret_cast<const float*>(p.slot_ptr(cur)) + p.origin_offset();
        P.vel = reinterpret_cast<const float*>(v.slot_ptr(0)) + v.origin_offset();
        P.out_sx = px->stride; P.out_sy = py->stride;
        P.v_sx = vx->stride; P.v_sy = vy->stride;
        P.nx = int(px->domain); P.ny = int(py->domain); P.nz = int(pz->domain);
        P.x_begin = int(box.b[0]); P.x_end = int(box.e[0]);
        P.y_begin = int(box.b[1]); P.y_end = int(box.e[1]);
        P.z_begin = int(box.b[2]); P.z_end = int(box.e[2]);
        P.pad_x = int(px->pad_l); P.pad_y = int(py->pad_l); P.pad_z = int(pz->pad_l);
        P.vpad_x = int(vx->pad_l); P.vpad_y = int(vy->pad_l); P.vpad_z = int(vz->pad_l);
        P.pol_c = pol_c; P.pol_h = pol_h; P.pol_pv = pol_pv; P.st_cs = st_cs; P.peer_probe = peer_probe;
        for (int r = 0; r <= ISO_MAX_R; r++) P.c[r] = r <= radius ? float(coef[r]) : 0.f;
    }

    int launch(Solution& s, int, int64_t t, const Box& box, cudaStream_t st) override {
        if (box.empty()) return 0;
        const Var& p = s.vars[0];
        const int cur = p.slot_of(t);
        IsoParams P{};
        fill_params(s, cur, box, P);
        int mode = s.fp_mode;
        bool use_tma = maps_ok && kernel != "direct";
        if (kernel == "auto") {
            // thin slabs in y/z (halo faces) are not worth a tile sweep
            if (box.e[1] - box.b[1] < 8 || box.e[2] - box.b[2] < 16) use_tma = false;
        }
        if (kernel == "tma" && !maps_ok) return set_error(YB_EUNSUPPORTED, "the tiled TMA kernel needs sm_90+");
        if (use_tma) {
            const int ti = radius == 8 ? tile : 0;      // other radii: one compiled variant
            const TileCfg& c = radius == 8 ? tile_cfg(tile) : *iso_radius_cfg(radius);
            if (mem_probe && c.fn[3]) mode = 3;
            // fused halo exchange: only for whole-domain launches of a kernel that implements the peer stores
            P.peer_lo = P.peer_hi = nullptr;
            bool fused = s.fused_x.var == 0 && c.fused_ok && mode != 3 && box.b[0] == 0 && box.e[0] == P.nx && P.nx >= 2 * radius;
            // in-kernel completion signal: needs room for boundary-first chunks; the copy-engine path exists only with it
            const bool sig = fused && s.fused_x.counter != nullptr && P.nx >= 4 * radius;
            if (fused && s.fused_x.dma && !sig) fused = false;
            const bool nb_lo = fused && s.fused_x.lo != nullptr, nb_hi = fused && s.fused_x.hi != nullptr;   // x neighbours
            if (fused) {
                if (!s.fused_x.dma) {        // the kernel stores the boundary planes into the neighbours itself
                    P.peer_lo = static_cast<float*>(s.fused_x.lo);
                    P.peer_hi = static_cast<float*>(s.fused_x.hi);
                }
                s.fused_x.used = true;
            }
            P.nty = int((box.e[1] - box.b[1] + c.ty - 1) / c.ty);
            P.ntz = int((box.e[2] - box.b[2] + c.tz - 1) / c.tz);
            const int64_t ntile = int64_t(P.nty) * P.ntz;
            const int gmax = grid_override > 0 ? grid_override : num_sms;
            // x chunks of equal length.  With an in-kernel completion signal the chunk that starts at plane 0 and the chunk
            // that ends at plane nx-1 (swept downwards) are numbered first, so that every CTA computes -- and stores into the
            // neighbours -- the boundary planes at the very start of its first sweeps.
            const int64_t ib = box.b[0], ie = box.e[0];
            const int64_t nxb = ie - ib;
            const int min_nc = (sig && nb_lo && nb_hi) ? 2 : 1;
            const int64_t max_nc = std::min<int64_t>(ISO_MAX_CHUNKS, sig ? nxb / radius : nxb);
            int nc_best = min_nc;
            if (lx > 0) {
                nc_best = int(std::max<int64_t>(min_nc, std::min<int64_t>((nxb + lx - 1) / lx, max_nc)));
            } else {
                // Pick the chunk count that minimises (rounds of units per CTA) x (chunk length + queue
                // warm-up): long chunks amortise the 2R warm-up planes, short ones balance the last round.
                double best = 1e30;
                for (int64_t nc = min_nc; nc <= max_nc; nc++) {
                    const int64_t l = (nxb + nc - 1) / nc;
                    const int64_t rounds = (ntile * nc + gmax - 1) / gmax;
                    const double cost = double(rounds) * (double(l) + 2 * radius * 0.4);
                    if (cost < best * 0.999) { best = cost; nc_best = int(nc); }
                }
            }
            int nb = 0, nsig = 0;
            {
                const int nc = nc_best;      // balanced partition: chunk lengths differ by at most one plane (all >= R when signalling)
                std::vector<int> order;
                if (sig && nb_lo) order.push_back(0);
                if (sig && nb_hi && !(nc == 1 && nb_lo)) order.push_back(nc - 1);
                nsig = int(order.size());
                for (int c = 0; c < nc; c++)
                    if (std::find(order.begin(), order.end(), c) == order.end()) order.push_back(c);
                for (int c : order) {
                    const int64_t x0 = ib + nxb * c / nc, x1 = ib + nxb * (c + 1) / nc;
                    const bool down = sig && nb_hi && c == nc - 1 && !(nc == 1 && nb_lo);
                    P.cx0[nb] = int(down ? x1 - 1 : x0); P.clen[nb] = int(x1 - x0); P.cdir[nb] = down ? -1 : 1;
                    nb++;
                }
            }
            P.nchunks = nb;
            P.sig_units = 0; P.sig_total = 0; P.sig_counter = nullptr; P.sig_flag_lo = P.sig_flag_hi = nullptr; P.sig_epoch = 0;
            const int64_t nunits = int64_t(P.nty) * P.ntz * P.nchunks;
            int grid = int(std::min<int64_t>(nunits, gmax));
            if (sig) {
                P.sig_units = int(ntile) * nsig;
                P.sig_total = unsigned(grid) * unsigned(c.ty * c.tz / 8 / 32);     // consumer warps: 8 points per thread
                P.sig_counter = s.fused_x.counter;
                P.sig_flag_lo = s.fused_x.flag_lo; P.sig_flag_hi = s.fused_x.flag_hi;
                P.sig_epoch = s.fused_x.epoch;
                s.fused_x.signalled = true;
            }
            if (!attr_set[ti][mode]) {
                YB_CUDA(cudaFuncSetAttribute(c.fn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, int(c.smem)));
                attr_set[ti][mode] = true;
            }
            c.fn[mode]<<<grid, c.threads, c.smem, st>>>(maps[ti][cur], P);
        } else {
            dim3 blk(128, 1, 1);
            dim3 grd(unsigned((box.e[2] - box.b[2] + 127) / 128), unsigned(box.e[1] - box.b[1]), unsigned(box.e[0] - box.b[0]));
            if (grd.y > 65535 || grd.z > 65535) return set_error(YB_EUNSUPPORTED, "direct kernel: domain too large in x or y");
            switch (mode) {
                case 0: iso3dfd_direct_kernel<0><<<grd, blk, 0, st>>>(P, radius); break;
                case 1: iso3dfd_direct_kernel<1><<<grd, blk, 0, st>>>(P, radius); break;
                default: iso3dfd_direct_kernel<2><<<grd, blk, 0, st>>>(P, radius); break;
            }
        }
        YB_CUDA(cudaGetLastError());
        return 1;
    }

    // In-run tuner: the same variants as the offline tuner, one per step of a live run.
    static constexpr int TUNE_TILES[6] = {7, 6, 5, 4, 1, 0};
    static constexpr int TUNE_LXS[3] = {0, 256, 512};
    int tune_variants(const Solution&) const override { return (maps_ok && radius == 8 && kernel != "direct") ? 18 : 0; }
    void tune_select(Solution& s, int v) override {
        tile = TUNE_TILES[v / 3]; lx = TUNE_LXS[v % 3];
        s.options["tile"] = std::to_string(tile);
        s.options["lx"] = std::to_string(lx);
        preload_kernel((const void*)tile_cfg(tile).fn[s.fp_mode]);
    }
    std::string tune_describe(const Solution&, int v) const override {
        char b[160];
        snprintf(b, sizeof b, "tile=%d (%s) lx=%d", TUNE_TILES[v / 3], tile_cfg(TUNE_TILES[v / 3]).name, TUNE_LXS[v % 3]);
        return b;
    }

    // Offline tuner: the compiled sweep variants (tile shape, producer warpgroup, planes per trip) x sweep chunk
    // lengths, timed over the whole rank box; the analogue of the reference's block-size search
    // (/root/reference/src/kernel/lib/auto_tuner.cpp) for the knobs this engine has.
    int auto_tune(Solution& s, cudaStream_t st, std::string& report) override {
        if (!maps_ok || radius != 8 || kernel == "direct") { report = "iso3dfd: no tiled variants to tune for this configuration"; return 0; }
        Box whole;
        for (int d = 0; d < 3; d++) { whole.b[d] = 0; whole.e[d] = s.rank_size[d]; }
        if (whole.e[1] < 8 || whole.e[2] < 16) { report = "iso3dfd: domain too thin for the tiled kernel"; return 0; }
        const int64_t t0 = s.vars[0].last_valid_step();
        const int tiles[] = {7, 6, 5, 4, 1, 0};
        const int lxs[] = {0, 256, 512, 1024};   // planes per interior chunk (0 = cost model)
        const int keep_tile = tile, keep_lx = lx;
        double best = 1e30;
        int best_tile = tile, best_lx = lx;
        char line[160];
        report.clear();
        for (int ti : tiles)
            for (int l : lxs) {
                if (l > whole.e[0]) continue;
                tile = ti; lx = l;
                int64_t t = t0;
                const double ms = time_launches(st, 3, [&]() { return launch(s, 0, t++, whole, st); });
                if (ms < 0) { tile = keep_tile; lx = keep_lx; return set_error(YB_ECUDA, "auto-tuner: a trial launch failed"); }
                snprintf(line, sizeof line, " tile=%d (%s) lx=%d: %.4f ms/step\n", ti, tile_cfg(ti).name, l, ms);
                report += line;
                if (ms < best) { best = ms; best_tile = ti; best_lx = l; }
            }
        tile = best_tile; lx = best_lx;
        snprintf(line, sizeof line, "best: tile=%d lx=%d (%.4f ms/step)\n", tile, lx, best);
        report += line;
        s.options["tile"] = std::to_string(tile);
        s.options["lx"] = std::to_string(lx);
        return 0;
    }
};

}  // namespace

std::unique_ptr<Engine> make_iso3dfd_engine() { return std::unique_ptr<Engine>(new IsoEngine()); }

}  // namespace yb
