// comm.cuh -- multi-GPU plumbing shared by the kernels' host wrappers (see comm.cu).
#pragma once
#include <vector>
#include "common.cuh"
#include "scalars.cuh"

// Halo of a row shard: which owned entries go to which peer, and where each peer's entries
// land in the ghost tail of an [owned ; ghost] vector.
struct lsspg_halo {
    int n_owned = 0, n_ghost = 0, n_send = 0, npeers = 0;
    std::vector<int> peers, send_off, recv_off;
    int *d_send_idx = nullptr;     // [n_send] owned row indices, grouped by peer
    double *d_send_buf = nullptr;  // [n_send]
};

namespace lsspg {
inline bool distributed(const lsspg_ctx *ctx) { return ctx->comm != nullptr; }
int comm_allreduce(lsspg_ctx *ctx, double *d_buf, int count);
// all-reduce scal[slot, slot+K) across ranks, then run `fin` on the device
int red_post(lsspg_ctx *ctx, int slot, int K, const FinProg &fin, bool guarded);
int halo_exchange(lsspg_ctx *ctx, const lsspg_halo *H, double *dx);
}  // namespace lsspg
