// dist.cu -- reference-picture hand-off between the GPUs of one node, reachable from a C host.
//
// The path shards by picture, like the reference's frame threads (libovvc/ovdec.c:189-248): picture k of the decoding order is
// reconstructed by rank k mod G.  The only exchange is what rcn_inter_synchronization / ovdpb_synchro_ref_decoded_ctus
// (libovvc/rcn_inter.c:131-146, libovvc/dpb.c:1242-1269) wait for on the CPU: a finished reference picture must reach the ranks
// whose reference picture lists name it.  ovgpu_picture_publish (owner) / ovgpu_picture_consume (every other rank that needs it)
// move the three planes over NVLink with grouped ncclSend / ncclRecv on the calling context's stream; ordering against the
// kernels is stream order / events, never a host wait.  No collective is fused into a kernel because no kernel of this path
// produces partial results for one: the hand-off is a copy of a finished picture.
//
// NCCL is loaded at run time (dlopen): a host that never calls ovgpu_dist_* needs no NCCL, and a host that already carries one
// (torch's) shares it.
#include "common.cuh"
#include <dlfcn.h>
#include <stdlib.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0, ncclUint8_ = 1 };

struct NcclApi {
    void *lib;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(int);
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.lib) return 0;
    static const char *names[] = { "libnccl.so.2", "libnccl.so" };
    void *h = NULL;
    for (unsigned i = 0; i < 2 && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return -1;
    NcclApi a;
    a.lib = h;
#define SYM(field, name) *(void **)&a.field = dlsym(h, name); if (!a.field) { return -1; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl = a;
    return 0;
}

struct OVGPUDist {
    ncclComm_t comm;
    int nb_ranks, rank, device;
    uint64_t bytes_sent, bytes_received;
};

#define OVGPU_NCCL_CHECK(ctx, call)                                                                                  \
    do {                                                                                                             \
        const int r__ = (call);                                                                                      \
        if (r__ != ncclSuccess_) return ovgpu_set_error((ctx), OVGPU_ECUDA, "%s: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

extern "C" {

int ovgpu_dist_unique_id(void *id, size_t bytes)
{
    if (!id || bytes < sizeof(ncclUniqueId)) return OVGPU_EINVAL;
    if (nccl_load()) return OVGPU_ENODEV;
    return g_nccl.GetUniqueId((ncclUniqueId *)id) == ncclSuccess_ ? 0 : OVGPU_ECUDA;
}

int ovgpu_dist_create(OVGPUCtx *ctx, int nb_ranks, int rank, const void *id, OVGPUDist **out)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !out || !id || nb_ranks < 1 || rank < 0 || rank >= nb_ranks) return OVGPU_EINVAL;
    if (nccl_load()) return ovgpu_set_error(ctx, OVGPU_ENODEV, "NCCL (libnccl.so.2) cannot be loaded: %s", dlerror());
    OVGPUDist *d = (OVGPUDist *)calloc(1, sizeof(*d));
    if (!d) return OVGPU_ENOMEM;
    d->nb_ranks = nb_ranks; d->rank = rank; d->device = ctx->device;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const int r = g_nccl.CommInitRank(&d->comm, nb_ranks, uid, rank);
    if (r != ncclSuccess_) { free(d); return ovgpu_set_error(ctx, OVGPU_ECUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    *out = d;
    return 0;
}

void ovgpu_dist_destroy(OVGPUDist *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->comm) g_nccl.CommDestroy(d->comm);
    free(d);
}

int ovgpu_picture_publish(OVGPUCtx *ctx, OVGPUDist *d, const OVGPUPicture *pic, const int *consumers, int nb_consumers)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !d || !pic || nb_consumers < 0 || (nb_consumers && !consumers)) return OVGPU_EINVAL;
    for (int i = 0; i < nb_consumers; i++)
        if (consumers[i] < 0 || consumers[i] >= d->nb_ranks || consumers[i] == d->rank) return OVGPU_EINVAL;
    if (!nb_consumers) return 0;
    OVGPU_NCCL_CHECK(ctx, g_nccl.GroupStart());
    for (int i = 0; i < nb_consumers; i++)
        for (int c = 0; c < 3; c++) {
            const size_t bytes = pic->pitch[c] * (size_t)(c ? pic->height >> 1 : pic->height);
            OVGPU_NCCL_CHECK(ctx, g_nccl.Send(pic->plane[c], bytes, ncclUint8_, consumers[i], d->comm, ctx->stream));
            d->bytes_sent += bytes;
        }
    OVGPU_NCCL_CHECK(ctx, g_nccl.GroupEnd());
    return 0;
}

int ovgpu_picture_consume(OVGPUCtx *ctx, OVGPUDist *d, OVGPUPicture *pic, int owner)
{
    OVGPU_DEVICE_GUARD(ctx);
    if (!ctx || !d || !pic || owner < 0 || owner >= d->nb_ranks || owner == d->rank) return OVGPU_EINVAL;
    OVGPU_NCCL_CHECK(ctx, g_nccl.GroupStart());
    for (int c = 0; c < 3; c++) {
        const size_t bytes = pic->pitch[c] * (size_t)(c ? pic->height >> 1 : pic->height);
        OVGPU_NCCL_CHECK(ctx, g_nccl.Recv(pic->plane[c], bytes, ncclUint8_, owner, d->comm, ctx->stream));
        d->bytes_received += bytes;
    }
    OVGPU_NCCL_CHECK(ctx, g_nccl.GroupEnd());
    return 0;
}

int ovgpu_dist_traffic(const OVGPUDist *d, uint64_t *bytes_sent, uint64_t *bytes_received)
{
    if (!d) return OVGPU_EINVAL;
    if (bytes_sent) *bytes_sent = d->bytes_sent;
    if (bytes_received) *bytes_received = d->bytes_received;
    return 0;
}

} // extern "C"
