// iso3dfd point arithmetic, shared by every iso3dfd kernel (yb_iso3dfd.cuh, yb_iso3dfd_tt.cuh).
//
// Equation and association order: /root/reference/src/stencils/Iso3dfdStencil.cpp:63-137 (get_next_p):
//   acc = p*c0;  r=1..R: acc += (((((p[x-r]+p[x+r])+p[y-r])+p[y+r])+p[z-r])+p[z+r]) * c_r
//   p(t+1) = ((2*p) - p(t-1)) + acc*v
// FP modes: 0 strict IEEE mul/add, 1 canonical FMA, 2 = the FMA pattern GCC emits for the reference's default build.
//
// YB_DEVFN is `__device__ __forceinline__` in the product build.  The CTA emulator of the test suite
// (tests/emul/tt_emul.cpp) compiles this header with g++, defines YB_DEVFN as `static inline` and supplies the
// rounded-arithmetic intrinsics as plain C functions (built with -ffp-contract=off).
#pragma once
#ifndef YB_DEVFN
#define YB_DEVFN __device__ __forceinline__
#endif

namespace yb {

template <int MODE>
YB_DEVFN float iso_group(float acc, float pc, float c0, float cr, float xm, float xp, float ym, float yp,
                                           float zm, float zp, bool first) {
    float s = __fadd_rn(xm, xp);
    s = __fadd_rn(s, ym);
    s = __fadd_rn(s, yp);
    s = __fadd_rn(s, zm);
    s = __fadd_rn(s, zp);
    if (MODE == 0) {
        if (first) acc = __fmul_rn(pc, c0);
        return __fadd_rn(acc, __fmul_rn(s, cr));
    } else if (MODE == 1) {
        if (first) acc = __fmul_rn(pc, c0);
        return __fmaf_rn(s, cr, acc);
    } else {
        if (first) return __fmaf_rn(pc, c0, __fmul_rn(s, cr));
        return __fmaf_rn(s, cr, acc);
    }
}

template <int MODE>
YB_DEVFN float iso_final(float acc, float pc, float prev, float v) {
    // 2*p is exact, so fma(2,p,-prev) == round((2*p) - prev): one instruction, same bits.
    float lhs = __fmaf_rn(2.0f, pc, -prev);
    if (MODE == 0) return __fadd_rn(lhs, __fmul_rn(acc, v));
    return __fmaf_rn(acc, v, lhs);
}

}  // namespace yb
