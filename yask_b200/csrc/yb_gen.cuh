// Support for emitter-generated stencil kernels (yask_b200/emitter/yask_cuda_emit.py).
// Host side: the tables a generated file fills in (what the reference's generated context ctor encodes,
// /root/reference/src/compiler/lib/YaskKernel.cpp:730-).  Device side: the statement vocabulary
// (RD/WR/C/ADD/SUB/MUL/DIV) the generated kernels are written in.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace yb { namespace gen {

constexpr int GEN_MAX_ACC = 32;   // distinct (var, step-offset) pairs one part may touch
constexpr int GEN_BLOCK = 128;    // threads per CTA along the unit-stride dim

struct GenParams {
    int xb, xe, yb, ye, zb, ze;              // box to compute, rank-local domain coordinates
    void* ptr[GEN_MAX_ACC];                  // element (0,0,0) of the var's step slot for each access
    long long sx[GEN_MAX_ACC], sy[GEN_MAX_ACC], sz[GEN_MAX_ACC];   // element strides (0 where the var lacks the dim)
};

typedef void (*GenKernelFn)(const GenParams);

struct GenVar {
    const char* name;
    std::vector<const char*> dims;   // declared order, step dim first if any
    int alloc_t;
    bool is_output;
    int l1_norm;
    std::vector<int> halo_l, halo_r; // per solution domain dim
};
struct GenAccess { int var, toff; };
struct GenPart {
    const char* name;
    int fp_ops, reads, writes;
    std::vector<GenAccess> acc;
    std::vector<int> outs;           // indices into acc of the written accesses
    GenKernelFn fn[2][2];            // [fp64?][mode: 0 strict, 1 fused]
};
struct GenStage { const char* name; std::vector<GenPart> parts; };
struct GenStencil {
    std::string name, step_dim;
    int elem_bytes = 4;
    std::vector<std::string> domain_dims;
    std::vector<GenVar> vars;
    std::vector<GenStage> stages;
};

#define GEN_FN(k, T, M) reinterpret_cast<yb::gen::GenKernelFn>(static_cast<void (*)(const yb::gen::GenParams)>(k<T, M>))

#ifdef __CUDACC__
// MODE 0: every operation individually rounded, in statement order (== the reference built with
// -ffp-contract=off).  MODE 1: plain operators, nvcc may contract a*b+c into FMA (the analogue of the
// reference's default -ffp-contract=fast build; not bit-identical to GCC's choices).
template <typename T, int MODE> struct GenOp;
template <> struct GenOp<float, 0> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct GenOp<double, 0> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <typename T> struct GenOp<T, 1> {
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T mul(T a, T b) { return a * b; }
    static __device__ __forceinline__ T div(T a, T b) { return a / b; }
};

#define GEN_KERNEL_PROLOGUE                                            \
    const int z = P.zb + blockIdx.x * GEN_BLOCK + threadIdx.x;         \
    const int y = P.yb + blockIdx.y;                                   \
    const int x = P.xb + blockIdx.z;                                   \
    if (z >= P.ze) return;
#define RD(a, dx, dy, dz) (static_cast<const T*>(P.ptr[a])[(x + (dx)) * P.sx[a] + (y + (dy)) * P.sy[a] + (z + (dz)) * P.sz[a]])
#define WR(a, v) static_cast<T*>(P.ptr[a])[x * P.sx[a] + y * P.sy[a] + z * P.sz[a]] = (v)
#define C(v) static_cast<T>(v)
#define ADD(a, b) GenOp<T, MODE>::add(a, b)
#define SUB(a, b) GenOp<T, MODE>::sub(a, b)
#define MUL(a, b) GenOp<T, MODE>::mul(a, b)
#define DIV(a, b) GenOp<T, MODE>::div(a, b)
#endif

} }  // namespace yb::gen
