// dev.h -- device-code portability layer.
//
// Kernels are written as structs with phase functions (phase0 .. phaseN-1, separated by a CTA barrier)
// over an explicit (tid, block index) so the SAME source compiles two ways:
//   * nvcc, sm_100a: ovd_kernel<K> below is the __global__ entry; intrinsics map to the native
//     Blackwell packed-16-bit instructions (VIADDMNMX.S16x2, VIMNMX.S16x2, VIADD.16x2, IDP.2A, PRMT);
//   * g++ with -DOVGPU_HOST_EMU (tests/emu only): phases run in a serial loop over the CTA's threads and
//     intrinsics are bit-exact C models. This is a TEST harness used to check kernel logic against the
//     oracle on machines without a GPU; the product library never contains or calls it.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef OVGPU_HOST_EMU
#include <stdlib.h>
#include <vector>
#define OVD static inline
#define OVHD static inline          /* helpers shared by host code and kernels */
#define OVDM inline                 /* non-static member functions (functors) */
#define OVD_CONST static const
#define OVD_CTAB static const
struct ovd_dim3 { unsigned x, y, z; ovd_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
#else
#include <cuda_runtime.h>
#define OVD static __device__ __forceinline__
#define OVHD static __host__ __device__ __forceinline__
#define OVDM __device__ __forceinline__
#define OVD_CONST __device__ const
#define OVD_CTAB __constant__ const       /* small tables indexed (nearly) uniformly across a warp: constant cache */
typedef dim3 ovd_dim3;
#endif

// ---- packed 2 x int16 arithmetic -------------------------------------------------------------------------
#ifdef OVGPU_HOST_EMU
static inline int16_t ovd__lo(uint32_t v) { return (int16_t)(v & 0xffff); }
static inline int16_t ovd__hi(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t ovd__pk(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }
static inline int ovd__mn(int a, int b) { return a < b ? a : b; }
static inline int ovd__mx(int a, int b) { return a > b ? a : b; }
// per half: min((int16)(a+b), c)   (the add wraps to 16 bit like the hardware)
OVD uint32_t ovd_viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c)
{ return ovd__pk(ovd__mn((int16_t)(ovd__lo(a) + ovd__lo(b)), ovd__lo(c)), ovd__mn((int16_t)(ovd__hi(a) + ovd__hi(b)), ovd__hi(c))); }
// per half: max(min((int16)(a+b), c), 0)
OVD uint32_t ovd_viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c)
{ return ovd__pk(ovd__mx(ovd__mn((int16_t)(ovd__lo(a) + ovd__lo(b)), ovd__lo(c)), 0), ovd__mx(ovd__mn((int16_t)(ovd__hi(a) + ovd__hi(b)), ovd__hi(c)), 0)); }
OVD uint32_t ovd_vmax_s16x2(uint32_t a, uint32_t b)
{ return ovd__pk(ovd__mx(ovd__lo(a), ovd__lo(b)), ovd__mx(ovd__hi(a), ovd__hi(b))); }
OVD uint32_t ovd_vmin_s16x2(uint32_t a, uint32_t b)
{ return ovd__pk(ovd__mn(ovd__lo(a), ovd__lo(b)), ovd__mn(ovd__hi(a), ovd__hi(b))); }
OVD uint32_t ovd_vadd_16x2(uint32_t a, uint32_t b)
{ return ovd__pk(ovd__lo(a) + ovd__lo(b), ovd__hi(a) + ovd__hi(b)); }
OVD uint32_t ovd_vsub_16x2(uint32_t a, uint32_t b)
{ return ovd__pk(ovd__lo(a) - ovd__lo(b), ovd__hi(a) - ovd__hi(b)); }
// a = 2 x s16, b = 4 x s8: LO: a.lo*b0 + a.hi*b1 + c ; HI: a.lo*b2 + a.hi*b3 + c
OVD int ovd_dp2a_lo(uint32_t a, uint32_t b, int c)
{ return c + ovd__lo(a) * (int8_t)(b & 0xff) + ovd__hi(a) * (int8_t)((b >> 8) & 0xff); }
OVD int ovd_dp2a_hi(uint32_t a, uint32_t b, int c)
{ return c + ovd__lo(a) * (int8_t)((b >> 16) & 0xff) + ovd__hi(a) * (int8_t)((b >> 24) & 0xff); }
// PTX prmt.b32, default mode: nibble bits 0-2 select a byte of {y,x}, bit 3 replicates that byte's sign
OVD uint32_t ovd_prmt(uint32_t x, uint32_t y, uint32_t s)
{
    uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 0xf;
        uint32_t b = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
// CUDA __byte_perm: 3-bit selectors only
OVD uint32_t ovd_byte_perm(uint32_t x, uint32_t y, uint32_t s) { return ovd_prmt(x, y, s & 0x7777u); }
OVD int ovd_sad(int a, int b, int c) { return c + (a > b ? a - b : b - a); }
// low 32 bits of (hi:lo) >> (sh & 31)   (SHF.R)
OVD uint32_t ovd_funnel_r(uint32_t lo, uint32_t hi, int sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
OVD int ovd_min(int a, int b) { return a < b ? a : b; }
OVD int ovd_max(int a, int b) { return a > b ? a : b; }
OVD uint32_t ovd_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
#else
OVD uint32_t ovd_viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }
OVD uint32_t ovd_viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }
OVD uint32_t ovd_vmax_s16x2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
OVD uint32_t ovd_vmin_s16x2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
OVD uint32_t ovd_vadd_16x2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
OVD uint32_t ovd_vsub_16x2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
OVD int ovd_dp2a_lo(uint32_t a, uint32_t b, int c) { return __dp2a_lo((int)a, (int)b, c); }
OVD int ovd_dp2a_hi(uint32_t a, uint32_t b, int c) { return __dp2a_hi((int)a, (int)b, c); }
OVD uint32_t ovd_byte_perm(uint32_t x, uint32_t y, uint32_t s) { return __byte_perm(x, y, s); }   // 3-bit selectors only
// prmt.b32 default mode: selector nibble bit 3 replicates the sign of the selected byte (PRMT)
OVD uint32_t ovd_prmt(uint32_t x, uint32_t y, uint32_t s)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(s));
    return r;
}
OVD int ovd_sad(int a, int b, int c) { return (int)__sad(a, b, (unsigned)c); }
OVD uint32_t ovd_funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, (unsigned)sh); }
OVD int ovd_min(int a, int b) { return min(a, b); }
OVD int ovd_max(int a, int b) { return max(a, b); }
OVD uint32_t ovd_umin(uint32_t a, uint32_t b) { return min(a, b); }
#endif

OVD int ovd_clamp(int v, int lo, int hi) { return ovd_min(ovd_max(v, lo), hi); }
// floor(log2(v)) for v >= 1, 0 for v == 0 (one FLO instead of a shift loop)
#ifdef OVGPU_HOST_EMU
OVD int ovd_floor_log2(unsigned v) { return v > 1 ? 31 - __builtin_clz(v) : 0; }
#else
OVD int ovd_floor_log2(unsigned v) { return v > 1 ? 31 - __clz((int)v) : 0; }
#endif

// ---- global memory access ----------------------------------------------------------------------------------
struct ovd_u2 { uint32_t x, y; };
struct ovd_u4 { uint32_t x, y, z, w; };

#ifdef OVGPU_HOST_EMU
OVD ovd_u2 ovd_ld8(const void *p) { ovd_u2 r; memcpy(&r, p, 8); return r; }
OVD ovd_u4 ovd_ld16(const void *p) { ovd_u4 r; memcpy(&r, p, 16); return r; }
OVD ovd_u4 ovd_ld16_rw(const void *p) { ovd_u4 r; memcpy(&r, p, 16); return r; }
OVD void ovd_st8(void *p, ovd_u2 v) { memcpy(p, &v, 8); }
OVD void ovd_st16(void *p, ovd_u4 v) { memcpy(p, &v, 16); }
#else
// streaming (touch-once) accesses: bypass L1 allocation.  ovd_ld8 / ovd_ld16 use the READ-ONLY path (ld.global.nc): only for
// memory that no thread of the running kernel writes; kernels that work in place use ovd_ld16_rw.
OVD ovd_u2 ovd_ld8(const void *p)
{
    ovd_u2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
OVD ovd_u4 ovd_ld16(const void *p)
{
    ovd_u4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
OVD ovd_u4 ovd_ld16_rw(const void *p)
{
    ovd_u4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
OVD void ovd_st8(void *p, ovd_u2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
OVD void ovd_st16(void *p, ovd_u4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
#endif

// ---- global atomics (kernels that build work lists on the device) -------------------------------------------------------------
#ifdef OVGPU_HOST_EMU
OVD void ovd_atomic_max(int *p, int v) { if (v > *p) *p = v; }
OVD void ovd_atomic_or(uint32_t *p, uint32_t v) { *p |= v; }
OVD void ovd_threadfence_system() {}
#else
OVD void ovd_threadfence_system() { __threadfence_system(); }      // orders stores to mapped host memory
OVD void ovd_atomic_max(int *p, int v) { atomicMax(p, v); }
OVD void ovd_atomic_or(uint32_t *p, uint32_t v) { atomicOr(p, v); }
#endif

// ---- asynchronous copies global -> shared (cp.async = LDGSTS): 16 bytes, both addresses 16-byte aligned, L1 bypassed ----------
#ifdef OVGPU_HOST_EMU
OVD void ovd_cp_async16(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
OVD void ovd_cp_async_wait_all() {}
#else
OVD void ovd_cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// the calling thread's copies have landed; a barrier afterwards publishes them to the other threads
OVD void ovd_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

// ---- kernel entry / emulation driver -----------------------------------------------------------------------
// A kernel struct K provides:  Args, State, THREADS, MIN_BLOCKS, PHASES and
//   static void phase<P>(const Args&, unsigned char *smem, State&, int tid, int bx, int by, int bz)
#ifndef OVGPU_HOST_EMU
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MIN_BLOCKS) ovd_kernel(const typename K::Args a)
{
    extern __shared__ __align__(16) unsigned char ovd_smem[];
    typename K::State st;
    const int tid = threadIdx.x, bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    K::template phase<0>(a, ovd_smem, st, tid, bx, by, bz);
    if (K::PHASES > 1) { __syncthreads(); K::template phase<1>(a, ovd_smem, st, tid, bx, by, bz); }
    if (K::PHASES > 2) { __syncthreads(); K::template phase<2>(a, ovd_smem, st, tid, bx, by, bz); }
    if (K::PHASES > 3) { __syncthreads(); K::template phase<3>(a, ovd_smem, st, tid, bx, by, bz); }
    if (K::PHASES > 4) { __syncthreads(); K::template phase<4>(a, ovd_smem, st, tid, bx, by, bz); }
    if (K::PHASES > 5) { __syncthreads(); K::template phase<5>(a, ovd_smem, st, tid, bx, by, bz); }
    if (K::PHASES > 6) { __syncthreads(); K::template phase<6>(a, ovd_smem, st, tid, bx, by, bz); }
    if (K::PHASES > 7) { __syncthreads(); K::template phase<7>(a, ovd_smem, st, tid, bx, by, bz); }
}

// K::L1_HEAVY (optional): shared-memory carve-out in percent for kernels whose global loads rely on L1 hits
template <class K> static inline auto ovd_l1_heavy(int) -> decltype(K::L1_HEAVY, 0) { return K::L1_HEAVY; }
template <class K> static inline int ovd_l1_heavy(long) { return 0; }

template <class K>
static inline cudaError_t ovd_launch(const typename K::Args &a, dim3 grid, size_t smem, cudaStream_t s)
{
    // per kernel instantiation AND device: function attributes are per device.  Two threads may both find the bit clear and
    // both set the (idempotent) attributes; the bit is published only after they are set.
    static unsigned configured_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned bit = 1u << (dev & 31);
    if (!(__atomic_load_n(&configured_mask, __ATOMIC_ACQUIRE) & bit)) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(ovd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        // ask for the largest shared-memory carve-out so that occupancy is bounded by registers, not by the L1 split
        // (kernels that lean on L1 for their global loads say so with K::L1_HEAVY and keep the default split)
        if (smem > 0 && !ovd_l1_heavy<K>(0)) cudaFuncSetAttribute(ovd_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        else if (ovd_l1_heavy<K>(0)) cudaFuncSetAttribute(ovd_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, ovd_l1_heavy<K>(0));
        __atomic_fetch_or(&configured_mask, bit, __ATOMIC_RELEASE);
    }
    ovd_kernel<K><<<grid, K::THREADS, smem, s>>>(a);
    return cudaGetLastError();
}
#else
// CTAs of a launch may run in any order on the GPU: tests run the emulation in both orders to catch kernels whose
// result depends on it (e.g. a store that spills into a neighbouring block's samples)
static int ovd_emu_reverse_blocks = 0;
template <class K>
static inline void ovd_emulate(const typename K::Args &a, ovd_dim3 grid, size_t smem_bytes)
{
    std::vector<typename K::State> st(K::THREADS);
    std::vector<uint64_t> smem((smem_bytes + 7) / 8 + 2);
    for (unsigned bz_ = 0; bz_ < grid.z; bz_++)
        for (unsigned by_ = 0; by_ < grid.y; by_++)
            for (unsigned bx_ = 0; bx_ < grid.x; bx_++) {
                const unsigned bx = ovd_emu_reverse_blocks ? grid.x - 1 - bx_ : bx_, by = ovd_emu_reverse_blocks ? grid.y - 1 - by_ : by_,
                               bz = ovd_emu_reverse_blocks ? grid.z - 1 - bz_ : bz_;
                unsigned char *sm = (unsigned char *)smem.data();
                memset(sm, 0xcd, smem_bytes);   // poison: catches reads of unwritten shared memory
                for (int t = 0; t < K::THREADS; t++) K::template phase<0>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 1) for (int t = 0; t < K::THREADS; t++) K::template phase<1>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 2) for (int t = 0; t < K::THREADS; t++) K::template phase<2>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 3) for (int t = 0; t < K::THREADS; t++) K::template phase<3>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 4) for (int t = 0; t < K::THREADS; t++) K::template phase<4>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 5) for (int t = 0; t < K::THREADS; t++) K::template phase<5>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 6) for (int t = 0; t < K::THREADS; t++) K::template phase<6>(a, sm, st[t], t, bx, by, bz);
                if (K::PHASES > 7) for (int t = 0; t < K::THREADS; t++) K::template phase<7>(a, sm, st[t], t, bx, by, bz);
            }
}
#endif
