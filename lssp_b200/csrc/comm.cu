// comm.cu -- multi-GPU plumbing for the row-sharded solve loop: one process per GPU,
// NCCL over NVLink / NVSwitch.
//
// The path shards by contiguous row blocks (rows_g = [g*ceil(n/P), ...), the block boundaries
// of the reference's lssp_mat_get_block_diag, src/matrix-utils.cxx:615,626-628).  Two exchange
// steps exist, nothing else:
//   * halo exchange of x before every SpMV: each rank packs the owned entries its neighbours
//     need (one gather kernel) and the ghost segments travel with grouped ncclSend/ncclRecv
//     straight into the tail of the receiver's x vector ([owned ; ghost] layout);
//   * all-reduce (sum) of the 1-8 doubles of each dot/norm group, in place on the device
//     scalar slab, followed by a one-thread kernel that derives alpha/beta/omega from the
//     GLOBAL sums with the reference's IEEE operations.
// Triangular sweeps stay local to a rank (block-Jacobi), as BASELINE.json prescribes.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already loaded by torch, else the
// system one), so the library has no link-time dependency on it.
#include <dlfcn.h>
#include <string.h>
#include <vector>
#include "blas1.cuh"
#include "comm.cuh"
#include "spmv.cuh"

namespace lsspg {

// ---- minimal NCCL surface (stable C ABI since 2.x) -----------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSum = 0 };
enum { ncclFloat64 = 8 };

static struct {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} nccl;

static int nccl_bind()
{
    if (nccl.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (nccl.h) break;
    }
    LSSPG_CHECK(nccl.h != nullptr, "comm: cannot load libnccl.so.2 (%s)", dlerror());
#define BIND(f) *(void **)(&nccl.f) = dlsym(nccl.h, "nccl" #f); LSSPG_CHECK(nccl.f != nullptr, "comm: nccl" #f " not found")
    BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(AllReduce); BIND(Send); BIND(Recv);
    BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
    return 0;
}

#define LSSPG_NCCL(call)                                                                   \
    do {                                                                                   \
        int r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                          \
            set_error("NCCL error %d (%s) in %s", r__, nccl.GetErrorString(r__), #call);   \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

constexpr int kP2PMaxRanks = 64;
constexpr int kP2PSlot = kMaxRedK + 1;   // K sums + the sequence number that announces them

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // one-shot peer-to-peer all-reduce (see k_p2p_allreduce_fin): this rank's mailbox and the peers' mailboxes,
    // mapped with CUDA IPC
    double *mail = nullptr;               // [2][kP2PMaxRanks][kP2PSlot]
    double **d_peer_mail = nullptr;       // device array of nranks mailbox pointers (own one included)
    std::vector<void *> opened;
    unsigned long long seq = 0;
};

int comm_allreduce(lsspg_ctx *ctx, double *d_buf, int count)
{
    Comm *c = (Comm *)ctx->comm;
    if (!c || c->nranks == 1) return 0;
    LSSPG_NCCL(nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, c->comm, ctx->stream));
    return 0;
}

__global__ void k_fin(FinProg fin, double *scal, int *flags, const int *stop)
{
    if (stop && *stop) return;
    fin_run(fin, scal, flags);
}

// One-shot all-reduce of K <= 8 doubles over NVLink, fused with the FinProg that derives alpha / beta / omega from the
// global sums: ONE launch of one CTA instead of ncclAllReduce (8 bytes: ~12 us of launch + protocol latency) followed by
// a one-thread kernel.  Every rank stores its partial sums and then the sequence number of this reduction into slot
// [seq & 1][rank] of EVERY rank's mailbox (peer stores through the IPC mappings), waits until its own mailbox shows the
// sequence number in all P slots, and adds the P contributions in rank order -- the same order on every rank, so all
// ranks hold bit-identical sums, run to run.  Two slot generations suffice: a rank can only be one reduction ahead of
// the slowest one, because finishing reduction s needs everybody's contribution to s.
__global__ void __launch_bounds__(64) k_p2p_allreduce_fin(double *const *peer_mail, int rank, int nranks, unsigned long long seq, int K,
                                                          int slot, FinProg fin, double *scal, int *flags, const int *stop, int *err)
{
    const int t = threadIdx.x;
    const size_t gen = (size_t)(seq & 1) * kP2PMaxRanks * kP2PSlot;
    if (t < nranks) {
        double *dst = peer_mail[t] + gen + (size_t)rank * kP2PSlot;
        for (int k = 0; k < K; k++) asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + k), "d"(scal[slot + k]) : "memory");
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + kMaxRedK), "l"(seq) : "memory");
        const double *src = peer_mail[rank] + gen + (size_t)t * kP2PSlot;
        unsigned long long got = 0;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src + kMaxRedK) : "memory");
            if (++spins > (1ll << 26)) { *err = 1; break; }
        } while (got != seq);
    }
    __syncthreads();
    if (t == 0) {
        const double *mine = peer_mail[rank] + gen;
        for (int k = 0; k < K; k++) {
            double s = 0.0;
            for (int r = 0; r < nranks; r++) {
                double v;
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(mine + (size_t)r * kP2PSlot + k) : "memory");
                s += v;
            }
            scal[slot + k] = s;
        }
        if (fin.n > 0 && !(stop && *stop)) fin_run(fin, scal, flags);
    }
}

// after a reducing kernel that deferred its FinProg: combine the ranks' partial sums, then derive
int red_post(lsspg_ctx *ctx, int slot, int K, const FinProg &fin, bool guarded)
{
    Comm *c = (Comm *)ctx->comm;
    if (c && c->nranks > 1 && c->d_peer_mail && K <= kMaxRedK) {
        c->seq++;
        LSSPG_LAUNCH(ctx, k_p2p_allreduce_fin, 1, 64, 0, c->d_peer_mail, c->rank, c->nranks, c->seq, K, slot, fin, ctx->d_scal,
                     ctx->d_flags, guarded ? ctx->d_flags + FLAG_STOP : (const int *)nullptr, ctx->d_flags + FLAG_TRI_TIMEOUT);
        return 0;
    }
    LSSPG_TRY(comm_allreduce(ctx, ctx->d_scal + slot, K));
    if (fin.n > 0)
        LSSPG_LAUNCH(ctx, k_fin, 1, 1, 0, fin, ctx->d_scal, ctx->d_flags,
                     guarded ? ctx->d_flags + FLAG_STOP : (const int *)nullptr);
    return 0;
}

__global__ void __launch_bounds__(kBlock) k_pack(int n, const int *__restrict__ idx, const double *__restrict__ x,
                                                 double *__restrict__ buf)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = x[idx[i]];
}

int halo_exchange(lsspg_ctx *ctx, const lsspg_halo *H, double *dx)
{
    Comm *c = (Comm *)ctx->comm;
    if (!H || H->npeers == 0) return 0;
    LSSPG_CHECK(c != nullptr, "halo_exchange: no communicator (call lsspg_comm_init first)");
    if (H->n_send > 0)
        LSSPG_LAUNCH(ctx, k_pack, stream_grid(ctx, H->n_send, kBlock), kBlock, 0, H->n_send, H->d_send_idx, dx, H->d_send_buf);
    LSSPG_NCCL(nccl.GroupStart());
    for (int p = 0; p < H->npeers; p++) {
        const int ns = H->send_off[p + 1] - H->send_off[p], nr = H->recv_off[p + 1] - H->recv_off[p];
        if (ns > 0) LSSPG_NCCL(nccl.Send(H->d_send_buf + H->send_off[p], (size_t)ns, ncclFloat64, H->peers[p], c->comm, ctx->stream));
        if (nr > 0) LSSPG_NCCL(nccl.Recv(dx + H->n_owned + H->recv_off[p], (size_t)nr, ncclFloat64, H->peers[p], c->comm, ctx->stream));
    }
    LSSPG_NCCL(nccl.GroupEnd());
    return 0;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

int lsspg_comm_unique_id(void *out128)
{
    LSSPG_TRY(nccl_bind());
    ncclUniqueId id;
    LSSPG_NCCL(nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return 0;
}

int lsspg_comm_init(lsspg_ctx *ctx, int rank, int nranks, const void *id128)
{
    LSSPG_CHECK(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "lsspg_comm_init: bad argument");
    LSSPG_TRY(nccl_bind());
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    Comm *c = new Comm();
    c->rank = rank; c->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    LSSPG_NCCL(nccl.CommInitRank(&c->comm, nranks, id, rank));
    ctx->comm = c;
    return 0;
}

// Peer-to-peer all-reduce, step 1: allocate this rank's mailbox and hand out its CUDA IPC handle (64 bytes), to be
// all-gathered by the launcher (torch.distributed in lssp_b200/dist.py).
int lsspg_comm_p2p_local(lsspg_ctx *ctx, void *handle64)
{
    Comm *c = (Comm *)ctx->comm;
    LSSPG_CHECK(c && handle64, "lsspg_comm_p2p_local: no communicator");
    LSSPG_CHECK(c->nranks <= kP2PMaxRanks, "lsspg_comm_p2p_local: more than %d ranks", kP2PMaxRanks);
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * 2 * kP2PMaxRanks * kP2PSlot;
    if (!c->mail) {
        LSSPG_CUDA(cudaMalloc(&c->mail, bytes));
        LSSPG_CUDA(cudaMemset(c->mail, 0, bytes));
    }
    cudaIpcMemHandle_t h;
    LSSPG_CUDA(cudaIpcGetMemHandle(&h, c->mail));
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

// step 2: map every peer's mailbox (handles[r] = rank r's 64-byte handle).  After this, dot products and norms are
// combined by k_p2p_allreduce_fin instead of ncclAllReduce.  Call on all ranks, after a barrier that follows step 1.
int lsspg_comm_p2p_connect(lsspg_ctx *ctx, const void *handles)
{
    Comm *c = (Comm *)ctx->comm;
    LSSPG_CHECK(c && c->mail && handles, "lsspg_comm_p2p_connect: call lsspg_comm_p2p_local first");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    std::vector<double *> ptrs(c->nranks, nullptr);
    for (int r = 0; r < c->nranks; r++) {
        if (r == c->rank) { ptrs[r] = c->mail; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, sizeof(h));
        void *p = nullptr;
        LSSPG_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->opened.push_back(p);
        ptrs[r] = (double *)p;
    }
    LSSPG_CUDA(cudaMalloc(&c->d_peer_mail, sizeof(double *) * c->nranks));
    LSSPG_CUDA(cudaMemcpy(c->d_peer_mail, ptrs.data(), sizeof(double *) * c->nranks, cudaMemcpyHostToDevice));
    return 0;
}

int lsspg_comm_destroy(lsspg_ctx *ctx)
{
    Comm *c = (Comm *)ctx->comm;
    if (!c) return 0;
    cudaStreamSynchronize(ctx->stream);
    for (void *p : c->opened) cudaIpcCloseMemHandle(p);
    cudaFree(c->d_peer_mail);
    cudaFree(c->mail);
    if (c->comm) nccl.CommDestroy(c->comm);
    delete c;
    ctx->comm = nullptr;
    return 0;
}

int lsspg_comm_size(lsspg_ctx *ctx, int *rank, int *nranks)
{
    Comm *c = (Comm *)ctx->comm;
    if (rank) *rank = c ? c->rank : 0;
    if (nranks) *nranks = c ? c->nranks : 1;
    return 0;
}

int lsspg_allreduce_sum(lsspg_ctx *ctx, double *d_buf, int count) { return comm_allreduce(ctx, d_buf, count); }

int lsspg_halo_create(lsspg_ctx *ctx, int n_owned, int npeers, const int *peers, const int *send_counts,
                      const int *h_send_idx, const int *recv_counts, lsspg_halo **out)
{
    LSSPG_CHECK(ctx && out && n_owned >= 0 && npeers >= 0, "lsspg_halo_create: bad argument");
    lsspg_halo *H = new lsspg_halo();
    H->n_owned = n_owned; H->npeers = npeers;
    H->send_off.assign(npeers + 1, 0);
    H->recv_off.assign(npeers + 1, 0);
    for (int p = 0; p < npeers; p++) {
        H->peers.push_back(peers[p]);
        H->send_off[p + 1] = H->send_off[p] + send_counts[p];
        H->recv_off[p + 1] = H->recv_off[p] + recv_counts[p];
    }
    H->n_send = H->send_off[npeers];
    H->n_ghost = H->recv_off[npeers];
    for (int i = 0; i < H->n_send; i++)
        LSSPG_CHECK(h_send_idx[i] >= 0 && h_send_idx[i] < n_owned, "lsspg_halo_create: send index %d outside the owned rows", h_send_idx[i]);
    LSSPG_CUDA(cudaMalloc(&H->d_send_idx, sizeof(int) * (size_t)(H->n_send > 0 ? H->n_send : 1)));
    LSSPG_CUDA(cudaMalloc(&H->d_send_buf, sizeof(double) * (size_t)(H->n_send > 0 ? H->n_send : 1)));
    if (H->n_send)
        LSSPG_CUDA(cudaMemcpyAsync(H->d_send_idx, h_send_idx, sizeof(int) * (size_t)H->n_send, cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = H;
    return 0;
}

int lsspg_halo_destroy(lsspg_ctx *ctx, lsspg_halo *H)
{
    if (!H) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(H->d_send_idx);
    cudaFree(H->d_send_buf);
    delete H;
    return 0;
}

int lsspg_halo_sizes(const lsspg_halo *H, int *n_owned, int *n_ghost, int *n_send)
{
    if (n_owned) *n_owned = H->n_owned;
    if (n_ghost) *n_ghost = H->n_ghost;
    if (n_send) *n_send = H->n_send;
    return 0;
}

int lsspg_halo_exchange(lsspg_ctx *ctx, const lsspg_halo *H, double *dx) { return halo_exchange(ctx, H, dx); }

int lsspg_csr_set_halo(lsspg_csr *A, lsspg_halo *H)
{
    LSSPG_CHECK(A != nullptr, "lsspg_csr_set_halo: NULL matrix");
    if (H) LSSPG_CHECK(A->num_rows == H->n_owned && A->num_cols == H->n_owned + H->n_ghost,
                       "lsspg_csr_set_halo: matrix is %d x %d but the halo describes %d owned + %d ghost columns",
                       A->num_rows, A->num_cols, H->n_owned, H->n_ghost);
    A->halo = H;
    return 0;
}

}  // extern "C"
