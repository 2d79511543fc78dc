// common.cuh -- shared host/device definitions of the B200 rcn backend (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/ovvc_b200.h"

struct OVGPUCtx {
    int device;
    cudaStream_t stream;
    int own_stream;
    int num_sms;
    uint64_t launches;
    char err[256];
    // profiling: events bracket each stage launch when enabled
    int profile;
    struct StageProf {
        double ms;
        uint64_t n;
    } prof[OVGPU_NB_STAGES];
    static const int MAX_PENDING = 256;
    struct Pending { cudaEvent_t a, b; int stage; } pending[MAX_PENDING];
    int nb_pending;
    cudaEvent_t ev_pool[2 * MAX_PENDING];
    int ev_pool_init;
    // scratch used to flush L2 between timed iterations
    void *flush_buf;
    size_t flush_bytes;
    // side streams for kernels of ONE stage that do not depend on each other (the six MC kernels work on disjoint blocks of the
    // picture): forked from / joined to `stream` with events, so that one kernel's tail overlaps the others instead of idling
    static const int NB_AUX = 5;
    cudaStream_t aux[NB_AUX];
    cudaEvent_t ev_fork, ev_join[NB_AUX];
    int aux_init;
};

// creates the side streams on first use; 0 on success
static inline int ovgpu_ctx_aux(OVGPUCtx *ctx)
{
    if (ctx->aux_init) return 0;
    for (int i = 0; i < OVGPUCtx::NB_AUX; i++)
        if (cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) != cudaSuccess) return -1;
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) return -1;
    ctx->aux_init = 1;
    return 0;
}

struct OVGPUPicture {
    int width, height, bitdepth;
    uint16_t *plane[3];
    size_t pitch[3];      // bytes
    int owned;            // planes allocated by us (single allocation based at plane[0])
    void *base;
    int device;
};

// Plane view passed by value to kernels: pitch in SAMPLES.
struct PlaneView {
    uint16_t *p;
    int pitch;   // samples
    int w, h;
};

static inline PlaneView plane_view(const OVGPUPicture *pic, int c)
{
    PlaneView v;
    v.p = pic->plane[c];
    v.pitch = (int)(pic->pitch[c] / sizeof(uint16_t));
    v.w = c ? pic->width >> 1 : pic->width;
    v.h = c ? pic->height >> 1 : pic->height;
    return v;
}

int ovgpu_set_error(OVGPUCtx *ctx, int code, const char *fmt, ...);

// Every C-ABI entry that allocates, copies or launches runs on the CONTEXT's device, whatever the calling thread's current
// device is (a decoder's entry threads default to device 0); the caller's device is restored on return.
struct DeviceGuard {
    int prev;
    explicit DeviceGuard(const OVGPUCtx *ctx) : prev(-1)
    {
        int cur = -1;
        if (ctx && cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device) { prev = cur; cudaSetDevice(ctx->device); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define OVGPU_DEVICE_GUARD(ctx) DeviceGuard ovgpu_device_guard__(ctx)

#define OVGPU_CUDA_CHECK(ctx, call)                                                          \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return ovgpu_set_error((ctx), OVGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__,   \
                                   #call, cudaGetErrorString(e__));                          \
    } while (0)

// Host-side section timer (debugging aid): accumulates wall time per labelled section of the commit / launch code, all
// contexts and threads together; ovgpu_host_timing_dump() prints the table.  Off (two loads per section) unless the
// environment has OVGPU_HOST_TIMING=1.
enum { HT_MC_LOOP, HT_MC_COPIES, HT_TU_CHECK, HT_TU_ITEMS, HT_TU_COPIES, HT_TU_ANCHORS, HT_INTRA_CHECK, HT_INTRA_LEVELS, HT_INTRA_COPIES,
       HT_BIND_LISTS, HT_BIND_COPIES, HT_RECON_MC, HT_RECON_LAUNCHES, HT_FENCE_WAIT, HT_NB };
extern unsigned long long ovgpu_ht_ns[HT_NB], ovgpu_ht_n[HT_NB];
extern int ovgpu_ht_on;
unsigned long long ovgpu_ht_now();
struct HostSection {
    int id; unsigned long long t0;
    explicit HostSection(int i) : id(i), t0(ovgpu_ht_on ? ovgpu_ht_now() : 0) {}
    ~HostSection() { if (ovgpu_ht_on) { __atomic_fetch_add(&ovgpu_ht_ns[id], ovgpu_ht_now() - t0, __ATOMIC_RELAXED); __atomic_fetch_add(&ovgpu_ht_n[id], 1ull, __ATOMIC_RELAXED); } }
};

// Stage bracket: counts the launch and, when profiling, records events around it.
struct StageScope {
    OVGPUCtx *ctx;
    int slot;
    StageScope(OVGPUCtx *c, int stage, int nb_kernels);
    ~StageScope();
};

// internal hooks between the stage files (reconstruction order, intra.cu <-> tr.cu)
struct ReconArgs;
int ovgpu_tu_bind_recon(OVGPUCtx *ctx, OVGPUTuBatch *b, uint64_t intra_serial, const uint32_t *pos_l, const uint32_t *pos_c,
                        int w_u, int h_u, const uint32_t *const order[2], const int nb[2]);
int ovgpu_tu_run_inter(OVGPUCtx *ctx, OVGPUPicture *pic, const OVGPUTuBatch *b, int plane);
void ovgpu_tu_recon_args(const OVGPUTuBatch *b, int plane, ReconArgs *a);
int ovgpu_tu_derive_scales(OVGPUCtx *ctx, const OVGPUPicture *pic, const OVGPUTuBatch *b, const uint16_t *d_wnd_bnd,
                           int min_idx, int max_idx_p1, int crs_offset);
int ovgpu_tu_nb_anchors(const OVGPUTuBatch *b);
int ovgpu_lmcs_scale_params(const OVGPULmcs *l, const uint16_t **d_wnd_bnd, int *min_idx, int *max_idx_p1, int *crs_offset);
int ovgpu_tu_check_picture(OVGPUCtx *ctx, const OVGPUPicture *pic, const OVGPUTuBatch *b);

#define OVGPU_CEIL_DIV(a, b) (((a) + (b) - 1) / (b))

// Fence on a pinned staging buffer: the host may rewrite the buffer only after the last H2D copy out of it has
// completed.  Waiting on this event (instead of the whole stream) lets the CPU stage picture N+1 while the GPU
// still works on picture N.
struct StagingFence {
    cudaEvent_t ev;
    int pending;
};
static inline cudaError_t fence_wait(StagingFence *f)
{
    cudaError_t e = cudaSuccess;
    if (f->pending) { HostSection hs(HT_FENCE_WAIT); e = cudaEventSynchronize(f->ev); f->pending = 0; }
    return e;
}
static inline cudaError_t fence_mark(StagingFence *f, cudaStream_t s)
{
    cudaError_t e = cudaSuccess;
    if (!f->ev) e = cudaEventCreateWithFlags(&f->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(f->ev, s);
    f->pending = e == cudaSuccess;
    return e;
}
static inline void fence_destroy(StagingFence *f) { if (f->ev) cudaEventDestroy(f->ev); f->ev = NULL; f->pending = 0; }

// ---- device helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// 16-byte streaming loads/stores: pictures are touched once per pass, keep them out of L1
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_u2(void *p, uint2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
#endif
