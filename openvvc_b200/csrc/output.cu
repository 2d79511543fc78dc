// output.cu -- output stage of a finished picture: conformance-window crop and the MD5 of the output byte stream, on the GPU.
//
// Reference behaviour: examples/dectest.c:376-406 (write_decoded_frame_to_file) writes the output window of every frame -- luma
// offsets are the window's chroma-unit offsets << 1 -- as Y rows, Cb rows, Cr rows of little-endian 16-bit samples, and the
// conformance script hashes that file (CI/ovnreg:82-85, md5sum over ALL frames).  ovgpu_picture_download_window copies exactly
// those rows to the host; ovgpu_md5_picture folds them into a running MD5 on the device instead, so that a conformance run moves
// 16 bytes per stream over PCIe rather than 25 MB per 4K picture.  MD5 is a serial chain (RFC 1321: 64 dependent steps per
// 64-byte block), so this is ONE thread per hash and about 0.2 s per 4K picture: a checking aid for the drop-in tests, on its own
// stream beside the reconstruction, not part of the timed path.
#include "common.cuh"
#include "pp_kernels.h"
#include "fg_kernels.h"
#include <stdlib.h>

struct OVGPUMd5 {
    uint32_t *d_state;       // a, b, c, d, length low / high (bytes), fill (bytes buffered), 16 words of buffer
    uint32_t *h_state;
    int device;
};

struct Md5Planes {
    const uint16_t *plane[3];
    uint32_t pitch[3], w[3], h[3];      // pitch in samples; cropped width / height in samples
};

__device__ __forceinline__ uint32_t md5_rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

__device__ void md5_block(uint32_t st[4], const uint32_t M[16])
{
    static const uint8_t S[64] = { 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                                   4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21 };
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be,
        0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
        0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
        0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
        0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
        0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391 };
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16)      { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d;          g = (3 * i + 5) & 15; }
        else             { f = c ^ (b | ~d);       g = (7 * i) & 15; }
        const uint32_t t = d;
        d = c; c = b;
        b = b + md5_rol(a + f + K[i] + M[g], S[i]);
        a = t;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

// one thread: folds the three cropped planes, row by row, into the running state
__global__ void ovgpu_md5_kernel(uint32_t *state, Md5Planes p)
{
    if (threadIdx.x || blockIdx.x) return;
    uint32_t st[4] = { state[0], state[1], state[2], state[3] };
    uint64_t len = (uint64_t)state[4] | ((uint64_t)state[5] << 32);
    uint32_t fill = state[6];                       // bytes in the buffer: even (samples are 2 bytes)
    uint32_t M[16];
#pragma unroll
    for (int i = 0; i < 16; i++) M[i] = state[7 + i];
    for (int c = 0; c < 3; c++)
        for (uint32_t y = 0; y < p.h[c]; y++) {
            const uint16_t *row = p.plane[c] + (size_t)y * p.pitch[c];
            uint32_t x = 0;
            // leading sample(s) until the buffer is word aligned and the source 4-byte aligned
            while (x < p.w[c] && ((fill & 2) || ((uintptr_t)(row + x) & 3))) {
                const uint32_t s = row[x++], wi = fill >> 2;
                // (dynamic index into M: kept in local memory only on this path)
                uint32_t v = (fill & 2) ? (M[wi] | (s << 16)) : s;
                M[wi] = v;
                fill += 2;
                if (fill == 64) { md5_block(st, M); fill = 0; }
            }
            // aligned middle: two samples per word
            for (; x + 2 <= p.w[c] && !(fill & 2); x += 2) {
                M[fill >> 2] = *(const uint32_t *)(row + x);
                fill += 4;
                if (fill == 64) { md5_block(st, M); fill = 0; }
            }
            for (; x < p.w[c]; x++) {
                const uint32_t s = row[x], wi = fill >> 2;
                M[wi] = (fill & 2) ? (M[wi] | (s << 16)) : s;
                fill += 2;
                if (fill == 64) { md5_block(st, M); fill = 0; }
            }
            len += (uint64_t)p.w[c] * 2;
        }
    state[0] = st[0]; state[1] = st[1]; state[2] = st[2]; state[3] = st[3];
    state[4] = (uint32_t)len; state[5] = (uint32_t)(len >> 32); state[6] = fill;
    for (int i = 0; i < 16; i++) state[7 + i] = M[i];
}

// padding + length (RFC 1321 section 3.1-3.2) and the digest bytes
__global__ void ovgpu_md5_final_kernel(uint32_t *state, uint32_t *digest)
{
    if (threadIdx.x || blockIdx.x) return;
    uint32_t st[4] = { state[0], state[1], state[2], state[3] };
    const uint64_t bits = ((uint64_t)state[4] | ((uint64_t)state[5] << 32)) * 8;
    uint32_t fill = state[6];
    uint32_t M[16];
    for (int i = 0; i < 16; i++) M[i] = i < (int)((fill + 3) >> 2) ? state[7 + i] : 0;
    if (fill & 2) M[fill >> 2] = (M[fill >> 2] & 0xffffu) | 0x00800000u; else M[fill >> 2] = 0x80u;
    if (fill >= 56) {
        md5_block(st, M);
        for (int i = 0; i < 16; i++) M[i] = 0;
    }
    M[14] = (uint32_t)bits; M[15] = (uint32_t)(bits >> 32);
    md5_block(st, M);
    for (int i = 0; i < 4; i++) digest[i] = st[i];
}

static int window_of(const OVGPUPicture *pic, const OVGPUWindow *win, int c, int *x0, int *y0, int *w, int *h)
{
    // dectest.c:382-391: chroma offsets are the window's, luma offsets twice that
    const int l = win ? win->offset_lft : 0, r = win ? win->offset_rgt : 0, a = win ? win->offset_abv : 0, b = win ? win->offset_blw : 0;
    if (l < 0 || r < 0 || a < 0 || b < 0) return -1;
    const int sh = c ? 0 : 1;
    *x0 = l << sh; *y0 = a << sh;
    *w = (c ? pic->width >> 1 : pic->width) - ((l + r) << sh);
    *h = (c ? pic->height >> 1 : pic->height) - ((a + b) << sh);
    return (*w > 0 && *h > 0) ? 0 : -1;
}

extern "C" {

int ovgpu_picture_download_window(OVGPUCtx *ctx, const OVGPUPicture *pic, const OVGPUWindow *win, void *const data[3], const size_t linesize[3])
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !pic || !data || !linesize) return OVGPU_EINVAL;
    for (int c = 0; c < 3; c++) {
        int x0, y0, w, h;
        if (window_of(pic, win, c, &x0, &y0, &w, &h) || linesize[c] < (size_t)w * 2) return ovgpu_set_error(ctx, OVGPU_EINVAL, "output window does not fit the picture");
        OVGPU_CUDA_CHECK(ctx, cudaMemcpy2DAsync(data[c], linesize[c], (const char *)pic->plane[c] + (size_t)y0 * pic->pitch[c] + (size_t)x0 * 2, pic->pitch[c],
                                                (size_t)w * 2, h, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return 0;
}

int ovgpu_md5_create(OVGPUCtx *ctx, OVGPUMd5 **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out) return OVGPU_EINVAL;
    OVGPUMd5 *m = (OVGPUMd5 *)calloc(1, sizeof(*m));
    if (!m) return OVGPU_ENOMEM;
    m->device = ctx->device;
    if (cudaMalloc(&m->d_state, 32 * sizeof(uint32_t)) != cudaSuccess || cudaMallocHost(&m->h_state, 32 * sizeof(uint32_t)) != cudaSuccess) {
        ovgpu_md5_destroy(m);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "MD5 state");
    }
    *out = m;
    return ovgpu_md5_reset(ctx, m);
}

void ovgpu_md5_destroy(OVGPUMd5 *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->d_state) cudaFree(m->d_state);
    if (m->h_state) cudaFreeHost(m->h_state);
    free(m);
}

int ovgpu_md5_reset(OVGPUCtx *ctx, OVGPUMd5 *m)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !m) return OVGPU_EINVAL;
    static const uint32_t init[4] = { 0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u };
    OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));         // (h_state may still be the source of an earlier reset)
    memset(m->h_state, 0, 32 * sizeof(uint32_t));
    memcpy(m->h_state, init, sizeof(init));
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_state, m->h_state, 32 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int ovgpu_md5_picture(OVGPUCtx *ctx, OVGPUMd5 *m, const OVGPUPicture *pic, const OVGPUWindow *win)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !m || !pic) return OVGPU_EINVAL;
    Md5Planes p;
    for (int c = 0; c < 3; c++) {
        int x0, y0, w, h;
        if (window_of(pic, win, c, &x0, &y0, &w, &h)) return ovgpu_set_error(ctx, OVGPU_EINVAL, "output window does not fit the picture");
        p.plane[c] = pic->plane[c] + (size_t)y0 * (pic->pitch[c] / 2) + x0;
        p.pitch[c] = (uint32_t)(pic->pitch[c] / 2); p.w[c] = (uint32_t)w; p.h[c] = (uint32_t)h;
    }
    ovgpu_md5_kernel<<<1, 32, 0, ctx->stream>>>(m->d_state, p);
    OVGPU_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->launches++;
    return 0;
}

int ovgpu_md5_final(OVGPUCtx *ctx, OVGPUMd5 *m, uint8_t digest[16])
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !m || !digest) return OVGPU_EINVAL;
    // (the kernel pads a COPY of the state: the running hash can go on -- a digest per picture and one over the stream)
    ovgpu_md5_final_kernel<<<1, 32, 0, ctx->stream>>>(m->d_state, m->d_state + 24);
    OVGPU_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->launches++;
    OVGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(m->h_state + 24, m->d_state + 24, 16, cudaMemcpyDeviceToHost, ctx->stream));
    OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(digest, m->h_state + 24, 16);
    return 0;
}


// Up-scaling of an output picture to the sequence's maximum size: what pp_process_frame does on the CPU for sequences with
// reference picture resampling (post_proc.c:117-128 -> pp_sample_rate_conv, pp_pic_scale.c:252-377).  Asynchronous.
int ovgpu_picture_upscale(OVGPUCtx *ctx, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUScalingWindow *win)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !dst || !src || dst == src) return OVGPU_EINVAL;
    if (dst->bitdepth != src->bitdepth) return ovgpu_set_error(ctx, OVGPU_EINVAL, "up-scaling: bit depths differ");
    UpscaleK::Args a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < 3; c++) {
        PlaneView d = plane_view(dst, c), s = plane_view((OVGPUPicture *)src, c);
        a.dst[c] = d.p; a.dst_pitch[c] = d.pitch; a.src[c] = s.p; a.src_pitch[c] = s.pitch;
    }
    a.dst_w = dst->width; a.dst_h = dst->height; a.src_w = src->width; a.src_h = src->height;
    a.maxv = (1 << 10) - 1;                               // pp_pic_scale.c is compiled for 10 bit only (:40)
    if (upscale_setup(&a, win ? win->left : 0, win ? win->right : 0, win ? win->top : 0, win ? win->bottom : 0,
                      win ? win->chroma_hor_col != 0 : 1, win ? win->chroma_ver_col != 0 : 0))
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "up-scaling: the target (%d x %d) is smaller than the source window (%d x %d)", dst->width, dst->height, src->width, src->height);
    dim3 grid(OVGPU_CEIL_DIV(a.dst_w, 32), OVGPU_CEIL_DIV(a.dst_h, 8), 3);
    OVGPU_CUDA_CHECK(ctx, ovd_launch<UpscaleK>(a, grid, 0, ctx->stream));
    ctx->launches += 1;
    return 0;
}


// Film grain synthesis on an output picture: what pp_process_frame does on the CPU for streams with a film grain characteristics
// SEI (post_proc.c:50-57, 100-103 -> fg_grain_apply_pic, pp_film_grain.c:814-977).  The synthesiser object owns the pattern data
// base (built once on the host, 676 KB in HBM) and the per-picture generator states.
struct OVGPUFilmGrainSynth {
    int device;
    int8_t *d_db;
    uint32_t *d_seeds;
    size_t seeds_cap;            // entries
};

int ovgpu_film_grain_create(OVGPUCtx *ctx, OVGPUFilmGrainSynth **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out) return OVGPU_EINVAL;
    OVGPUFilmGrainSynth *s = (OVGPUFilmGrainSynth *)calloc(1, sizeof(*s));
    int8_t *db = (int8_t *)malloc(FG_DB_BYTES);
    if (!s || !db) { free(s); free(db); return OVGPU_ENOMEM; }
    s->device = ctx->device;
    fg_build_database(db);
    if (cudaMalloc(&s->d_db, FG_DB_BYTES) != cudaSuccess ||
        cudaMemcpyAsync(s->d_db, db, FG_DB_BYTES, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        free(db); ovgpu_film_grain_destroy(s);
        return ovgpu_set_error(ctx, OVGPU_ENOMEM, "film grain data base: %s", cudaGetErrorString(cudaGetLastError()));
    }
    free(db);
    *out = s;
    return 0;
}

void ovgpu_film_grain_destroy(OVGPUFilmGrainSynth *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->d_db) cudaFree(s->d_db);
    if (s->d_seeds) cudaFree(s->d_seeds);
    free(s);
}

int ovgpu_film_grain_picture(OVGPUCtx *ctx, OVGPUFilmGrainSynth *s, OVGPUPicture *dst, const OVGPUPicture *src, const OVGPUFilmGrain *fg, int32_t poc)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !s || !dst || !src || !fg || dst == src) return OVGPU_EINVAL;
    if (dst->width != src->width || dst->height != src->height) return ovgpu_set_error(ctx, OVGPU_EINVAL, "film grain: picture sizes differ");
    if (src->bitdepth != 10 || dst->bitdepth != 10) return ovgpu_set_error(ctx, OVGPU_EINVAL, "film grain: the reference synthesises at 10 bit only (pp_film_grain.c:836)");
    if (fg->log2_scale_factor > 15) return ovgpu_set_error(ctx, OVGPU_EINVAL, "film grain: log2_scale_factor %d", fg->log2_scale_factor);
    if (!fg_size_ok(src->width, src->height))
        return ovgpu_set_error(ctx, OVGPU_EINVAL, "film grain: %d x %d leaves a block row / column of fewer than 8 samples, which the reference reads out of bounds", src->width, src->height);
    static const uint8_t colour_offset[3] = { 0, 85, 170 };          // COLOUR_OFFSET_LUMA / _CR / _CB in component order (:56-58, 834-836)
    FilmGrainK::Args a;
    memset(&a, 0, sizeof(a));
    FgSeedK::Args sa;
    memset(&sa, 0, sizeof(sa));
    size_t need = 0;
    for (int c = 0; c < 3; c++) {
        const int cw = src->width >> (c != 0), chh = src->height >> (c != 0);
        sa.nb16w[c] = (cw + 15) >> 4; sa.nb16h[c] = (chh + 15) >> 4;
        need += (size_t)sa.nb16w[c] * sa.nb16h[c];
    }
    if (need > s->seeds_cap) {
        // (the buffer of a smaller picture may still be read by a kernel of the previous call on this stream)
        OVGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        if (s->d_seeds) cudaFree(s->d_seeds);
        s->d_seeds = NULL; s->seeds_cap = 0;
        if (cudaMalloc(&s->d_seeds, need * sizeof(uint32_t)) != cudaSuccess) return ovgpu_set_error(ctx, OVGPU_ENOMEM, "film grain generator states");
        s->seeds_cap = need;
    }
    size_t off = 0;
    for (int c = 0; c < 3; c++) {
        PlaneView d = plane_view(dst, c), sv = plane_view((OVGPUPicture *)src, c);
        a.dst[c] = d.p; a.dst_pitch[c] = d.pitch; a.src[c] = sv.p; a.src_pitch[c] = sv.pitch;
        sa.seeds[c] = s->d_seeds + off; a.seeds[c] = sa.seeds[c];
        off += (size_t)sa.nb16w[c] * sa.nb16h[c];
        sa.seed0[c] = fg_seeds[((uint32_t)poc + colour_offset[c]) % 256];
    }
    a.w = src->width; a.h = src->height; a.db = s->d_db;
    fg_make_params(&a.p, fg);
    OVGPU_CUDA_CHECK(ctx, ovd_launch<FgSeedK>(sa, dim3(1, 1, 3), 0, ctx->stream));
    dim3 grid(OVGPU_CEIL_DIV(a.w, FilmGrainK::COLS), OVGPU_CEIL_DIV(a.h, 16), 3);
    OVGPU_CUDA_CHECK(ctx, ovd_launch<FilmGrainK>(a, grid, FilmGrainK::SMEM_BYTES, ctx->stream));
    ctx->launches += 2;
    return 0;
}

} // extern "C"
