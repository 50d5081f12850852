// tri_tiled.cu -- tile-scheduled sparse triangular solve for structured-grid factors.
//
// The slice schedule of tri.cu pays one trip through L2 (~1-2 us) per dependency level;
// an ILU(0) factor of an N^3 7-point grid has 3N-2 levels, so its sweep is bound by that
// latency.  When the factor's off-diagonal offsets reveal a lattice (1, nx, nx*ny), rows are
// grouped into small boxes (8x8x8 by default).  One warp owns one box: dependencies INSIDE
// the box are served from the warp's shared-memory copy of x after a __syncwarp() (~100
// cycles per local level), only dependencies on OTHER boxes are polled in global memory
// with the sentinel protocol of tri.cu.  Seven of eight hops of the critical path thereby
// stay on chip.  Boxes are drawn as tickets in a topological order of the box graph (which
// is verified to be acyclic on the host), so the no-deadlock argument of tri.cu carries over.
//
// Arithmetic is untouched: every row still subtracts its products sequentially in the
// reference's order and divides by its stored diagonal -> bit-identical results
// (reference src/solver-tri.cxx:13-23, :35-45).
//
// HBM layout: the factor is stored in box-major row order as plain CSR (ptr/col/val) with
// perm/diag per slot; a box's rows are contiguous, so every sector is consumed completely
// while the box is resident.  Algorithmic bytes per sweep: 12 nnz(T) + 20 n.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <numeric>
#include "blas1.cuh"
#include "tri.cuh"
#include "host_par.h"

namespace lsspg {

constexpr unsigned long long kSentinelBitsT = 0xFFF8DEADBEEF0001ull;   // same value as tri.cu
constexpr int kMaxTileRows = 1024;

// Histogram of |column - row| over the off-diagonal entries (at most 16 distinct offsets: stencil factors have
// a handful).  Pieces of rows are counted by the host threads and merged; counts are order-independent.
struct OffsetHist {
    int m = 0;
    long long off[17], cnt[17];
    bool add(long long d, long long c)
    {
        for (int q = 0; q < m; q++)
            if (off[q] == d) { cnt[q] += c; return true; }
        if (m == 16) return false;
        off[m] = d; cnt[m] = c; m++;
        return true;
    }
};

bool detect_lattice(int n, const int *Tp, const int *Tj, int dims[3], std::vector<long long> &offsets)
{
    const int np = host_threads();
    std::vector<OffsetHist> part(np);
    std::vector<char> over(np, 0);
    parallel_ranges(n, [&](long long r0, long long r1, int p) {
        OffsetHist h;
        long long last = -1;
        int lastq = 0;
        for (int i = (int)r0; i < (int)r1; i++) {
            for (int k = Tp[i]; k < Tp[i + 1]; k++) {
                const long long d = llabs((long long)Tj[k] - i);
                if (d == 0) continue;
                if (d == last) { h.cnt[lastq]++; continue; }
                if (!h.add(d, 1)) { over[p] = 1; return; }
                last = d;
                for (lastq = 0; h.off[lastq] != d; lastq++) {}
            }
        }
        part[p] = h;
    }, np);
    std::map<long long, long long> hist;
    for (int p = 0; p < np; p++) {
        if (over[p]) return false;
        for (int q = 0; q < part[p].m; q++) hist[part[p].off[q]] += part[p].cnt[q];
    }
    if (hist.size() > 16) return false;
    offsets.clear();
    for (auto &kv : hist) offsets.push_back(kv.first);
    if (hist.size() < 2 || hist.begin()->first != 1) return false;
    const long long s3 = hist.rbegin()->first;
    if (s3 <= 1 || n % s3 != 0) return false;
    if (hist[1] < n / 2 || hist[s3] < n / 4) return false;
    long long s2 = 0, best = 0;
    for (auto &kv : hist) {
        if (kv.first > 1 && kv.first < s3 && s3 % kv.first == 0 && kv.second > best) {
            s2 = kv.first;
            best = kv.second;
        }
    }
    if (s2 && best >= n / 4) {
        dims[0] = (int)s2; dims[1] = (int)(s3 / s2); dims[2] = (int)(n / s3);
    }
    else {
        dims[0] = (int)s3; dims[1] = (int)(n / s3); dims[2] = 1;
    }
    return (long long)dims[0] * dims[1] * dims[2] == n;
}

// Set-up of the box schedule.  Only the dependency-level recurrence is inherently serial (one O(nnz) pass);
// everything else runs over the host threads on disjoint outputs, so the image does not depend on their number
// (tests/test_setup_threads.py pins it).
int tri_tiled_build_host(int which, int n, const int *Tp, const int *Tj, const double *Tx, TiledHost &H)
{
    const bool lower = (which == LSSPG_TRI_LOWER);
    int g[3];
    std::vector<long long> offsets;
    if (n < 4096 || !detect_lattice(n, Tp, Tj, g, offsets)) return 2;
    int t[3] = {8, 8, 8};
    if (g[2] == 1) { t[0] = 16; t[1] = 16; t[2] = 1; }
    bool shape_given = false;
    if (const char *e = getenv("LSSPG_TRI_TILE")) {
        int a, b, c;
        if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0) { t[0] = a; t[1] = b; t[2] = c; shape_given = true; }
    }
    if ((long long)t[0] * t[1] * t[2] > kMaxTileRows) return 2;
    // Skewed boxes (default; LSSPG_TRI_SKEW=0 keeps the plain grid).  An entry at offset d = dz nx ny + dy nx + dx couples grid point p with
    // p - (dx, dy, dz) (lower factor) or p + (dx, dy, dz) (upper).  With fill (ILU(1): offsets nx - 1, nx ny - nx,
    // nx ny - 1) some dx or dy are negative, neighbouring axis-aligned boxes then need each other and the box graph
    // is cyclic.  Boxes cut along u = x + s1 y + t1 z, v = y + s2 z, w = z instead, with the smallest s1, s2, t1 >= 0
    // that make u, v, w non-decreasing along every offset, are only ever coupled one way: the graph is acyclic by
    // construction (and verified below like any other).  s1 = s2 = t1 = 0 is the plain box grid.
    int sk[3] = {0, 0, 0};   // s1, s2, t1
    {
        const char *e = getenv("LSSPG_TRI_SKEW");
        if (!(e && atoi(e) == 0)) {
            struct V { long long x, y, z; };
            std::vector<V> vs;
            for (long long d : offsets) {
                long long dz = d / ((long long)g[0] * g[1]), r = d % ((long long)g[0] * g[1]);
                long long dy = r / g[0], dx = r % g[0];
                if (dx > g[0] / 2) { dx -= g[0]; dy += 1; }
                if (dy > g[1] / 2 && g[2] > 1) { dy -= g[1]; dz += 1; }
                vs.push_back({dx, dy, dz});
            }
            auto ceil_div = [](long long a, long long b) { return (a + b - 1) / b; };   // a >= 0, b > 0
            long long s1 = 0, s2 = 0, t1 = 0;
            for (const V &v : vs) {
                if (v.z == 0 && v.y > 0 && v.x < 0) s1 = std::max(s1, ceil_div(-v.x, v.y));
                if (v.z > 0 && v.y < 0) s2 = std::max(s2, ceil_div(-v.y, v.z));
            }
            for (const V &v : vs)
                if (v.z > 0 && v.x + s1 * v.y < 0) t1 = std::max(t1, ceil_div(-(v.x + s1 * v.y), v.z));
            if (s1 <= 4 && s2 <= 4 && t1 <= 8) { sk[0] = (int)s1; sk[1] = (int)s2; sk[2] = (int)t1; }
        }
    }
    // skewed 3-D boxes: u spans (1 + s1 + t1) nx, so boxes longer in u and flatter in w shorten the chain of boxes;
    // measured for ILUK(1) at 256^3 (profiles/r01_skew_tiles.log): 12x8x5 3.88 ms, 8x8x8 4.12, 16x8x4 4.14, 8x4x8 4.46
    if (!shape_given && g[2] > 1 && (sk[0] || sk[1] || sk[2])) { t[0] = 12; t[1] = 8; t[2] = 5; }
    // (rows wider than the unrolled in-box code paths, ILU(2) and up, take the generic loop of the kernel: verified
    // bit-exact on a B200 in round 2, profiles/r02_experimental_variants.txt)
    const long long ext[3] = {(long long)g[0] + (long long)sk[0] * (g[1] - 1) + (long long)sk[2] * (g[2] - 1),
                              (long long)g[1] + (long long)sk[1] * (g[2] - 1), g[2]};
    const long long ntl[3] = {(ext[0] + t[0] - 1) / t[0], (ext[1] + t[1] - 1) / t[1], (ext[2] + t[2] - 1) / t[2]};
    if (ntl[0] * ntl[1] * ntl[2] > (1ll << 28)) return 2;
    const int nt[3] = {(int)ntl[0], (int)ntl[1], (int)ntl[2]};
    int ntiles = nt[0] * nt[1] * nt[2];
    IVec tile_of((size_t)n);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            const int x = i % g[0], y = (i / g[0]) % g[1], z = i / (g[0] * g[1]);
            const int u = x + sk[0] * y + sk[2] * z, v = y + sk[1] * z;
            tile_of[i] = ((z / t[2]) * nt[1] + (v / t[1])) * nt[0] + (u / t[0]);
        }
    });
    if (sk[0] || sk[1] || sk[2]) {
        // skewed boxes that lie outside the grid are empty: renumber the others, order kept
        std::vector<int> newid((size_t)ntiles, 0);
        for (int i = 0; i < n; i++) newid[tile_of[i]] = 1;
        int cnt = 0;
        for (int k = 0; k < ntiles; k++) {
            const int has = newid[k];
            newid[k] = cnt;
            cnt += has;
        }
        parallel_ranges(n, [&](long long r0, long long r1, int) {
            for (int i = (int)r0; i < (int)r1; i++) tile_of[i] = newid[tile_of[i]];
        });
        ntiles = cnt;
    }
    // validate the triangle (threads) ...
    {
        const int np = host_threads();
        std::vector<char> bad(np, 0);
        parallel_ranges(n, [&](long long r0, long long r1, int p) {
            for (int i = (int)r0; i < (int)r1; i++) {
                const int b = Tp[i], e = Tp[i + 1];
                if (e <= b) { bad[p] = 1; return; }
                const int dpos = lower ? e - 1 : b;
                if (Tj[dpos] != i) { bad[p] = 1; return; }
                for (int k = b; k < e; k++) {
                    if (k == dpos) continue;
                    const int c = Tj[k];
                    if (lower ? !(c >= 0 && c < i) : !(c > i && c < n)) { bad[p] = 1; return; }
                }
            }
        }, np);
        for (char b : bad)
            if (b) return 2;
    }
    // ... and compute the dependency level of every row, in solve order.  Rows of a box are later grouped by
    // this GLOBAL level: the unfinished row of smallest level anywhere is then always inside its box's current
    // group with all operands produced, so level barriers inside boxes can never deadlock boxes that depend
    // on each other.
    IVec gl((size_t)n);
    int nlev_global = 0;
    for (int q = 0; q < n; q++) {
        const int i = lower ? q : n - 1 - q;
        const int b = lower ? Tp[i] : Tp[i] + 1, e = lower ? Tp[i + 1] - 1 : Tp[i + 1];
        int gg = 0;
        for (int k = b; k < e; k++) gg = std::max(gg, gl[Tj[k]] + 1);
        gl[i] = gg;
        nlev_global = std::max(nlev_global, gg + 1);
    }
    // box dependency edges (from * ntiles + to), each piece of rows deduplicated on its own, then merged
    std::vector<long long> edges;
    {
        const int np = host_threads();
        std::vector<std::vector<long long>> part(np);
        parallel_ranges(n, [&](long long r0, long long r1, int p) {
            std::vector<long long> &ev = part[p];
            long long last = -1;
            for (int i = (int)r0; i < (int)r1; i++) {
                const int ti = tile_of[i];
                for (int k = Tp[i]; k < Tp[i + 1]; k++) {
                    const int c = Tj[k];
                    if (c == i || tile_of[c] == ti) continue;
                    const long long ed = (long long)tile_of[c] * ntiles + ti;
                    if (ed != last) { ev.push_back(ed); last = ed; }
                }
            }
            std::sort(ev.begin(), ev.end());
            ev.erase(std::unique(ev.begin(), ev.end()), ev.end());
        }, np);
        for (auto &ev : part) edges.insert(edges.end(), ev.begin(), ev.end());
    }
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    // Box graph (producer box -> consumer box).  It may contain cycles (e.g. ILU(1) fill couples
    // x-neighbouring boxes both ways): boxes of one strongly connected component must simply be in
    // flight together, which the ticket order guarantees as long as a component is smaller than
    // the number of resident warps.  Tarjan (iterative) emits components consumers-first.
    std::vector<int> estart(ntiles + 1, 0);
    for (long long ed : edges) estart[ed / ntiles + 1]++;
    for (int k = 0; k < ntiles; k++) estart[k + 1] += estart[k];
    std::vector<int> idx(ntiles, -1), low(ntiles, 0), comp(ntiles, -1), stk, cu, ce;
    std::vector<char> onstk(ntiles, 0);
    int counter = 0, ncomp = 0, max_scc = 0;
    for (int root = 0; root < ntiles; root++) {
        if (idx[root] >= 0) continue;
        idx[root] = low[root] = counter++;
        stk.push_back(root); onstk[root] = 1;
        cu.push_back(root); ce.push_back(estart[root]);
        while (!cu.empty()) {
            const int u = cu.back();
            if (ce.back() < estart[u + 1]) {
                const int v = (int)(edges[ce.back()++] % ntiles);
                if (idx[v] < 0) {
                    idx[v] = low[v] = counter++;
                    stk.push_back(v); onstk[v] = 1;
                    cu.push_back(v); ce.push_back(estart[v]);
                }
                else if (onstk[v]) low[u] = std::min(low[u], idx[v]);
            }
            else {
                if (low[u] == idx[u]) {
                    int size = 0, w;
                    do { w = stk.back(); stk.pop_back(); onstk[w] = 0; comp[w] = ncomp; size++; } while (w != u);
                    max_scc = std::max(max_scc, size);
                    ncomp++;
                }
                cu.pop_back(); ce.pop_back();
                if (!cu.empty()) low[cu.back()] = std::min(low[cu.back()], low[u]);
            }
        }
    }
    if (max_scc > 128) return 2;
    // Boxes that depend on each other both ways could only be run with per-operand polling, which measured slower
    // than the slice schedule (round 1; that kernel is gone): such factors take the slice schedule.
    if (max_scc > 1) return 2;
    // producers-first rank of every component and its longest-path level in the condensation
    std::vector<int> rank_of(ntiles), tlc(ncomp, 0), byrank(ntiles), rstart(ncomp + 1, 0);
    for (int k = 0; k < ntiles; k++) { rank_of[k] = ncomp - 1 - comp[k]; rstart[rank_of[k] + 1]++; }
    for (int r = 0; r < ncomp; r++) rstart[r + 1] += rstart[r];
    {
        std::vector<int> pos(rstart.begin(), rstart.end() - 1);
        for (int k = 0; k < ntiles; k++) byrank[pos[rank_of[k]]++] = k;
    }
    for (int q = 0; q < ntiles; q++) {
        const int u = byrank[q];
        for (int k = estart[u]; k < estart[u + 1]; k++) {
            const int v = (int)(edges[k] % ntiles);
            if (rank_of[v] != rank_of[u]) tlc[rank_of[v]] = std::max(tlc[rank_of[v]], tlc[rank_of[u]] + 1);
        }
    }
    std::vector<int> tl(ntiles);
    for (int k = 0; k < ntiles; k++) tl[k] = tlc[rank_of[k]];
    // ticket order of the boxes: by component level, then component, then id (components stay contiguous)
    std::vector<int> torder(ntiles);
    std::iota(torder.begin(), torder.end(), 0);
    std::stable_sort(torder.begin(), torder.end(), [&](int a, int b) {
        return tl[a] != tl[b] ? tl[a] < tl[b] : rank_of[a] < rank_of[b];
    });
    std::vector<int> ticket_of(ntiles);
    for (int k = 0; k < ntiles; k++) ticket_of[torder[k]] = k;
    // slots: rows sorted by (ticket of box, local level, row)
    H.perm.resize((size_t)n);
    IVec &order = H.perm;
    H.tile_ptr.assign(ntiles + 1, 0);
    std::vector<int> &start = H.tile_ptr;
    {
        for (int i = 0; i < n; i++) start[ticket_of[tile_of[i]] + 1]++;
        for (int k = 0; k < ntiles; k++) start[k + 1] += start[k];
        std::vector<int> pos(start.begin(), start.end() - 1);
        for (int i = 0; i < n; i++) order[pos[ticket_of[tile_of[i]]]++] = i;   // bucket by box, rows ascending
    }
    H.max_tile_rows = 0;
    for (int k = 0; k < ntiles; k++) H.max_tile_rows = std::max(H.max_tile_rows, start[k + 1] - start[k]);
    if (H.max_tile_rows > kMaxTileRows) return 2;
    // per box: rows stably sorted by level; number of level groups and of off-diagonal entries
    IVec slot_of((size_t)n);
    H.ptr.resize((size_t)n + 1);
    H.lev_off.assign(ntiles + 1, 0);
    parallel_ranges(ntiles, [&](long long k0, long long k1, int) {
        for (int k = (int)k0; k < (int)k1; k++) {
            std::stable_sort(order.begin() + start[k], order.begin() + start[k + 1],
                             [&](int a, int b) { return gl[a] < gl[b]; });
            int groups = 0, cur = -1;
            for (int s = start[k]; s < start[k + 1]; s++) {
                const int i = order[s];
                slot_of[i] = s;
                H.ptr[s] = Tp[i + 1] - Tp[i] - 1;
                if (gl[i] != cur) { groups++; cur = gl[i]; }
            }
            H.lev_off[k] = groups + 1;   // group starts + the closing entry
        }
    }, 0, 64);
    const long long nent_total = parallel_exclusive_scan(H.ptr.data(), n);
    H.ptr[n] = (int)nent_total;
    {
        int run = 0;
        for (int k = 0; k < ntiles; k++) { const int c = H.lev_off[k]; H.lev_off[k] = run; run += c; }
        H.lev_off[ntiles] = run;
        H.lev_ptr.assign((size_t)run, 0);
    }
    H.n = n; H.which = which; H.num_tiles = ntiles; H.num_levels = nlev_global;
    H.num_tile_levels = ntiles ? *std::max_element(tl.begin(), tl.end()) + 1 : 0;
    H.acyclic = (max_scc <= 1);
    {   // predecessor boxes of every box, as tickets
        H.pred_ptr.assign(ntiles + 1, 0);
        for (long long ed : edges) H.pred_ptr[ticket_of[ed % ntiles] + 1]++;
        for (int k = 0; k < ntiles; k++) H.pred_ptr[k + 1] += H.pred_ptr[k];
        H.pred.resize(edges.size());
        std::vector<int> pos(H.pred_ptr.begin(), H.pred_ptr.end() - 1);
        for (long long ed : edges) H.pred[pos[ticket_of[ed % ntiles]]++] = ticket_of[ed / ntiles];
    }
    for (int k = 0; k < 3; k++) { H.tile_dims[k] = t[k]; H.grid_dims[k] = g[k]; }
    H.diag.resize((size_t)n);
    H.offdiag_nnz = (long long)Tp[n] - n;
    if (nent_total != H.offdiag_nnz) return 2;
    H.col.resize((size_t)H.offdiag_nnz);
    H.val.resize((size_t)H.offdiag_nnz);
    parallel_ranges(ntiles, [&](long long k0, long long k1, int) {
        for (int k = (int)k0; k < (int)k1; k++) {
            const int s0 = start[k], s1 = start[k + 1];
            int lp = H.lev_off[k], cur = -1;
            for (int s = s0; s < s1; s++) {
                const int i = order[s];
                if (gl[i] != cur) { H.lev_ptr[lp++] = s; cur = gl[i]; }
                const int b = Tp[i], e = Tp[i + 1];
                int ent = H.ptr[s];
                H.diag[s] = lower ? Tx[e - 1] : Tx[b];
                // application order: lower ascending storage order, upper descending
                for (int q = 0; q < e - b - 1; q++) {
                    const int kk = lower ? b + q : e - 1 - q;
                    const int c = Tj[kk];
                    H.col[ent] = (tile_of[c] == tile_of[i]) ? -(slot_of[c] - s0 + 1) : c;
                    H.val[ent] = Tx[kk];
                    ent++;
                }
            }
            H.lev_ptr[lp] = s1;
        }
    }, 0, 64);
    return 0;
}

// ---- device side ---------------------------------------------------------------------
// Every box is packed into one 16-byte aligned blob
//   [lev: nlev+1 int][ptr: nrows+1 int][perm: nrows int][ext: next int][pred: npred int][diag: nrows f64][col: nent int][val: nent f64]
// (sections padded to 16 B; lev/ptr are box-relative; col < 0: -(slot in box + 1), col >= 0: index
// into ext, the list of rows of OTHER boxes this box reads) so that a single bulk async copy
// (cp.async.bulk, completion on an mbarrier) brings the whole box into shared memory.
struct BoxDesc {
    long long off;   // byte offset of the blob
    int bytes;       // blob size (multiple of 16)
    int nrows, nent, nlev, next, npred;
};

static inline size_t pad16(size_t b) { return (b + 15) & ~(size_t)15; }

struct BoxLayout {
    size_t lev, ptr, perm, ext, pred, diag, col, val, total;
};

__host__ __device__ inline BoxLayout box_layout(int nrows, int nent, int nlev, int next, int npred)
{
    BoxLayout L;
    size_t o = 0;
    L.lev = o;  o += (((size_t)(nlev + 1) * 4) + 15) & ~(size_t)15;
    L.ptr = o;  o += (((size_t)(nrows + 1) * 4) + 15) & ~(size_t)15;
    L.perm = o; o += (((size_t)nrows * 4) + 15) & ~(size_t)15;
    L.ext = o;  o += (((size_t)next * 4) + 15) & ~(size_t)15;
    L.pred = o; o += (((size_t)npred * 4) + 15) & ~(size_t)15;
    L.diag = o; o += (((size_t)nrows * 8) + 15) & ~(size_t)15;
    L.col = o;  o += (((size_t)nent * 4) + 15) & ~(size_t)15;
    L.val = o;  o += (((size_t)nent * 8) + 15) & ~(size_t)15;
    L.total = o;
    return L;
}

struct TiledArgs {
    const unsigned char *blob;
    const BoxDesc *desc;
    unsigned int *counter;
    int num_tiles;
    int blob_cap;        // shared-memory bytes reserved for one blob
    int max_tile_rows;
    int max_ext;
    double *x;
    const double *rhs;
    const int *stop;
    int *err;
    unsigned int *flags;        // per-box completion epochs (acyclic box graphs), else NULL
    unsigned int epoch;
    unsigned long long *prof;   // optional phase timers (LSSPG_TRI_PROF=1): ticket, blob, gather, compute, poll, boxes
};

__device__ __forceinline__ double ldx_relaxed(const double *p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void stx_relaxed(double *p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

// ---- acyclic box graphs: lean kernel -----------------------------------------------------
// When no two boxes depend on each other (ILU(0)-type factors) a box simply waits for the
// completion flags of the boxes it reads from; after that every operand it needs exists, so
// the group loop carries no polling and no sentinel tests.  The box is stored as ELL (fixed
// width w, column-major) whose padding entries multiply a private +0.0 operand by +0.0 --
// r - (+0) == r bit for bit -- which removes every predicate from the inner loop.  One warp
// executes ~50 instructions per group instead of ~350.
//   blob: [lev: nlev+1 int][perm: nrows int][ext: next int][pred: npred int][diag: nrows f64]
//         [ecol: w*nrows int][eval: w*nrows f64]
//   sx  : [0,nrows) rhs -> x of the box, [nrows, nrows+next) operands of other boxes, then +0.0
struct EllLayout {
    size_t lev, perm, ext, pred, diag, ecol, eval, total;
};

__host__ __device__ inline EllLayout ell_layout(int nrows, int w, int nlev, int next, int npred)
{
    EllLayout L;
    size_t o = 0;
    L.lev = o;  o += (((size_t)(nlev + 1) * 4) + 15) & ~(size_t)15;
    L.perm = o; o += (((size_t)nrows * 4) + 15) & ~(size_t)15;
    L.ext = o;  o += (((size_t)next * 4) + 15) & ~(size_t)15;
    L.pred = o; o += (((size_t)npred * 4) + 15) & ~(size_t)15;
    L.diag = o; o += (((size_t)nrows * 8) + 15) & ~(size_t)15;
    L.ecol = o; o += (((size_t)w * nrows * 4) + 15) & ~(size_t)15;
    L.eval = o; o += (((size_t)w * nrows * 8) + 15) & ~(size_t)15;
    L.total = o;
    return L;
}

// The dependency levels of one box, ELL width W known at compile time.  Only the loads of sx[col]
// depend on earlier levels; the row's columns, values, divisor and right-hand side do not, so they
// are fetched one pass AHEAD.  A pass = up to 32*R rows of ONE level, R rows per lane side by side
// (rows of a level are independent: their fp64 chains overlap), so a level of <= 64 rows costs one
// chain  LDS sx[col] -> W products -> W dependent subtractions (-> divide) -> STS -> __syncwarp.
template <int W, int R>
__device__ __forceinline__ void box_levels(const int *__restrict__ slev, int nlev, int nrows, int dummy,
                                           const int *__restrict__ ecol, const double *__restrict__ eval,
                                           const double *__restrict__ sdiag, double *sx, int lane)
{
    // Lanes without a row in the pass work on row 0's data and store into the spare slot `dummy`:
    // no predicates anywhere in the pass.
    int L = 0, base = slev[0];
    int st[R], c[R][W];
    double v[R][W], dg[R], r[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int slot = base + j * 32 + lane;
        const bool act = slot < slev[1];
        const int ls = act ? slot : 0;
        st[j] = act ? slot : dummy;
#pragma unroll
        for (int k = 0; k < W; k++) {
            c[j][k] = ecol[k * nrows + ls];
            v[j][k] = eval[k * nrows + ls];
        }
        dg[j] = sdiag[ls];
        r[j] = sx[ls];
    }
    for (;;) {
        int nL = L, nbase = base + 32 * R;
        bool newlevel = false;
        if (nbase >= slev[L + 1]) {
            nL = L + 1;
            newlevel = true;
            nbase = (nL < nlev) ? slev[nL] : 0;
        }
        const bool more = nL < nlev;
        const int nend = more ? slev[nL + 1] : 0;
        int nst[R], nc[R][W];
        double nv[R][W], ndg[R], nr[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int nslot = nbase + j * 32 + lane;
            const bool nact = nslot < nend;
            const int ls = nact ? nslot : 0;
            nst[j] = nact ? nslot : dummy;
#pragma unroll
            for (int k = 0; k < W; k++) {
                nc[j][k] = ecol[k * nrows + ls];
                nv[j][k] = eval[k * nrows + ls];
            }
            ndg[j] = sdiag[ls];
            nr[j] = sx[ls];   // rows of later passes still hold their right-hand side (row 0: any value, unused)
        }
        double xv[R][W];
#pragma unroll
        for (int j = 0; j < R; j++)
#pragma unroll
            for (int k = 0; k < W; k++) xv[j][k] = sx[c[j][k]];
#pragma unroll
        for (int k = 0; k < W; k++)
#pragma unroll
            for (int j = 0; j < R; j++) r[j] = r[j] - v[j][k] * xv[j][k];   // reference src/solver-tri.cxx:18 / :40
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (dg[j] != 1.0) r[j] = r[j] / dg[j];                         // :22 / :44 (x / 1.0 == x exactly)
            sx[st[j]] = r[j];
        }
        if (!more) break;
        if (newlevel) __syncwarp();
        L = nL;
        base = nbase;
#pragma unroll
        for (int j = 0; j < R; j++) {
            st[j] = nst[j];
            dg[j] = ndg[j];
            r[j] = nr[j];
#pragma unroll
            for (int k = 0; k < W; k++) {
                c[j][k] = nc[j][k];
                v[j][k] = nv[j][k];
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32, 8) tri_box_ell_kernel(const TiledArgs a)   // smem allows <= 7 boxes per SM anyway
{
    extern __shared__ __align__(128) unsigned char smem[];
    if (a.stop && *a.stop) return;
    const int lane = threadIdx.x;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
    unsigned char *sblob = smem + 16;
    double *sx = reinterpret_cast<double *>(sblob + a.blob_cap);
    const unsigned int bar_s = smem_u32(bar), blob_s = smem_u32(sblob);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned int phase = 0;
    const unsigned int total = (unsigned int)a.num_tiles + gridDim.x;
    for (;;) {
        unsigned int tk = 0;
        if (lane == 0) tk = atomicInc(a.counter, total - 1);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= (unsigned int)a.num_tiles) break;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        if (a.prof) t0 = clock64();
        const BoxDesc d = a.desc[tk];   // d.nent holds the ELL width here
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"((unsigned int)d.bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(blob_s), "l"(a.blob + d.off), "r"((unsigned int)d.bytes), "r"(bar_s) : "memory");
        }
        {
            unsigned int ok = 0;
            while (!ok) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(bar_s), "r"(phase) : "memory");
            }
            phase ^= 1;
        }
        const int nrows = d.nrows, w = d.nent;
        const EllLayout lay = ell_layout(nrows, w, d.nlev, d.next, d.npred);
        const int *slev = reinterpret_cast<const int *>(sblob + lay.lev);
        const int *sperm = reinterpret_cast<const int *>(sblob + lay.perm);
        const int *sext = reinterpret_cast<const int *>(sblob + lay.ext);
        const int *spred = reinterpret_cast<const int *>(sblob + lay.pred);
        const double *sdiag = reinterpret_cast<const double *>(sblob + lay.diag);
        const int *ecol = reinterpret_cast<const int *>(sblob + lay.ecol);
        const double *eval = reinterpret_cast<const double *>(sblob + lay.eval);
        // right-hand side of the box, eight independent loads in flight per lane
        for (int s0 = 0; s0 < nrows; s0 += 256) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int s = s0 + u * 32 + lane;
                v[u] = (s < nrows) ? __ldg(a.rhs + sperm[s]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int s = s0 + u * 32 + lane;
                if (s < nrows) sx[s] = v[u];
            }
        }
        if (lane == 0) sx[nrows + d.next] = 0.0;
        if (a.prof) t1 = clock64();
        // Hand-off between boxes without flags or fences: x is pre-filled with the sentinel NaN, so
        // every 8-byte operand announces itself.  (1) light gate: one lane per predecessor box polls
        // ONE operand of that box (the one it stores last) -- a handful of pollers per box, no storm;
        // (2) then all operands are fetched in one round and any that is still the sentinel (stores
        // of a box may land out of order) is re-polled on its own.
        for (int q = lane; q < d.npred; q += 32) {
            const double *f = a.x + spred[q];
            int spins = 0;
            while ((unsigned long long)__double_as_longlong(ldx_relaxed(f)) == kSentinelBitsT) {
                if (++spins > 16) __nanosleep(40);
                if (spins > (1 << 21)) { *a.err = 1; break; }
            }
        }
        __syncwarp();
        if (a.prof) t2 = clock64();
        for (int q0 = 0; q0 < d.next; q0 += 256) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int q = q0 + u * 32 + lane;
                v[u] = (q < d.next) ? ldx_relaxed(a.x + sext[q]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int q = q0 + u * 32 + lane;
                if (q < d.next) {
                    int spins = 0;
                    while ((unsigned long long)__double_as_longlong(v[u]) == kSentinelBitsT) {
                        v[u] = ldx_relaxed(a.x + sext[q]);
                        if (++spins > (1 << 21)) { *a.err = 1; break; }
                    }
                    sx[nrows + q] = v[u];
                }
            }
        }
        __syncwarp();
        if (a.prof) t3 = clock64();
        switch (w) {   // fixed widths: fully unrolled and software-pipelined (7-point ILU(0): 3, 5-point: 2)
            case 1: box_levels<1, 2>(slev, d.nlev, nrows, nrows + d.next + 1, ecol, eval, sdiag, sx, lane); break;
            case 2: box_levels<2, 2>(slev, d.nlev, nrows, nrows + d.next + 1, ecol, eval, sdiag, sx, lane); break;
            case 3: box_levels<3, 2>(slev, d.nlev, nrows, nrows + d.next + 1, ecol, eval, sdiag, sx, lane); break;
            case 4: box_levels<4, 2>(slev, d.nlev, nrows, nrows + d.next + 1, ecol, eval, sdiag, sx, lane); break;
            case 5: box_levels<5, 2>(slev, d.nlev, nrows, nrows + d.next + 1, ecol, eval, sdiag, sx, lane); break;   // ILU(1)
            case 6: box_levels<6, 2>(slev, d.nlev, nrows, nrows + d.next + 1, ecol, eval, sdiag, sx, lane); break;   // fill
            default:
                for (int L = 0; L < d.nlev; L++) {
                    const int sa = slev[L], sb = slev[L + 1];
                    for (int slot = sa + lane; slot < sb; slot += 32) {
                        double r = sx[slot];
                        for (int k = 0; k < w; k++) {
                            const int c = ecol[k * nrows + slot];
                            r = r - eval[k * nrows + slot] * sx[c];   // reference src/solver-tri.cxx:18 / :40
                        }
                        const double dg = sdiag[slot];
                        if (dg != 1.0) r = r / dg;                    // :22 / :44 (x / 1.0 == x exactly)
                        sx[slot] = r;
                    }
                    __syncwarp();
                }
        }
        if (a.prof) t4 = clock64();
        // publish: the values are their own ready flags (see the hand-off above): no fence, no flag store
        for (int s = lane; s < nrows; s += 32)
            asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(a.x + sperm[s]), "d"(sx[s]) : "memory");
        if (a.prof && lane == 0) {   // LSSPG_TRI_PROF=1: cycles per box by phase
            const long long t5 = clock64();
            atomicAdd(a.prof + 0, (unsigned long long)(t1 - t0));   // descriptor, blob, rhs gather
            atomicAdd(a.prof + 1, (unsigned long long)(t2 - t1));   // waiting for predecessor boxes
            atomicAdd(a.prof + 2, (unsigned long long)(t3 - t2));   // operands of other boxes
            atomicAdd(a.prof + 3, (unsigned long long)(t4 - t3));   // in-box levels
            atomicAdd(a.prof + 4, (unsigned long long)(t5 - t4));   // publish
            atomicAdd(a.prof + 5, 1ull);
        }
    }
}

int tri_tiled_solve(lsspg_ctx *ctx, const lsspg_tri *T, double *dx, const double *drhs, bool guarded)
{
    double sentinel;
    const unsigned long long bits = kSentinelBitsT;
    memcpy(&sentinel, &bits, sizeof(double));
    LSSPG_TRY(vec_set(ctx, T->n, dx, sentinel, guarded));
    TiledArgs a;
    a.blob = T->t_blob; a.desc = (const BoxDesc *)T->t_desc;
    a.counter = T->d_counter; a.num_tiles = T->num_tiles; a.blob_cap = T->blob_cap; a.max_tile_rows = T->max_tile_rows;
    a.max_ext = T->max_ext;
    a.flags = nullptr;
    a.epoch = ++const_cast<lsspg_tri *>(T)->epoch;
    a.x = dx; a.rhs = drhs;
    a.stop = guarded ? ctx->d_flags + FLAG_STOP : nullptr;
    a.err = ctx->d_flags + FLAG_TRI_TIMEOUT;
    const size_t smem = 16 + (size_t)T->blob_cap + 8 * ((size_t)T->max_tile_rows + T->max_ext + 2);
    static int env_cap = -1;
    if (env_cap < 0) {
        const char *e = getenv("LSSPG_TRI_TILED_CTAS_PER_SM");
        env_cap = e ? atoi(e) : 32;
        if (env_cap < 1) env_cap = 1;
    }
    int per_sm = (int)std::min<size_t>((size_t)env_cap, (size_t)(224 * 1024) / (smem + 1024));
    per_sm = std::max(1, std::min(per_sm, 32));
    int grid = std::min(T->num_tiles, ctx->num_sms * per_sm);
    if (grid < 1) grid = 1;
    static int prof_on = -1;
    static unsigned long long *d_prof = nullptr;
    if (prof_on < 0) {
        prof_on = getenv("LSSPG_TRI_PROF") ? 1 : 0;
        if (prof_on) { cudaMalloc(&d_prof, 64); cudaMemset(d_prof, 0, 64); }
    }
    a.prof = prof_on ? d_prof : nullptr;
    LSSPG_LAUNCH(ctx, tri_box_ell_kernel, grid, 32, smem, a);
    if (prof_on) {
        unsigned long long h[8];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, d_prof, 64, cudaMemcpyDeviceToHost);
        cudaMemset(d_prof, 0, 64);
        if (h[5])
            fprintf(stderr, "[tri_box_ell] boxes=%llu grid=%d cycles/box: fetch %.0f wait %.0f operands %.0f levels %.0f publish %.0f\n",
                    h[5], grid, (double)h[0] / h[5], (double)h[1] / h[5], (double)h[2] / h[5], (double)h[3] / h[5], (double)h[4] / h[5]);
    }
    return 0;
}

static int upload_common(lsspg_ctx *ctx, const TiledHost &H, lsspg_tri *T, const PackedBoxes &P)
{
    const auto &blob = P.blob;
    const std::vector<unsigned char> &desc = P.desc_bytes;
    const size_t cap = P.cap;
    const int max_ext = P.max_ext;
    T->tiled = true;
    T->num_tiles = H.num_tiles; T->max_tile_rows = H.max_tile_rows; T->num_tile_levels = H.num_tile_levels;
    T->blob_cap = (int)cap;
    T->max_ext = max_ext;
    T->box_flags = true;
    LSSPG_CUDA(cudaMalloc(&T->t_flags, sizeof(unsigned int) * std::max(H.num_tiles, 1)));
    LSSPG_CUDA(cudaMemsetAsync(T->t_flags, 0, sizeof(unsigned int) * std::max(H.num_tiles, 1), ctx->stream));
    for (int k = 0; k < 3; k++) { T->tile_dims[k] = H.tile_dims[k]; T->grid_dims[k] = H.grid_dims[k]; }
    LSSPG_CUDA(cudaMalloc(&T->t_blob, blob.size()));
    LSSPG_CUDA(cudaMalloc(&T->t_desc, std::max<size_t>(desc.size(), sizeof(BoxDesc))));
    LSSPG_CUDA(cudaMemcpyAsync(T->t_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!desc.empty())
        LSSPG_CUDA(cudaMemcpyAsync(T->t_desc, desc.data(), desc.size(), cudaMemcpyHostToDevice, ctx->stream));
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(tri_box_ell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// acyclic box graph: ELL blobs for tri_box_ell_kernel.  Host only; 2 = not applicable (a box would not fit).
// Boxes are sized, laid out and filled by the host threads (a box's blob is written by exactly one thread).
static int pack_ell_host(const TiledHost &H, PackedBoxes &P)
{
    const int nb = H.num_tiles;
    std::vector<BoxDesc> desc(nb);
    IVec pos_of_row((size_t)H.n);   // row -> position in box-major order
    parallel_ranges(H.n, [&](long long a, long long b, int) {
        for (int s = (int)a; s < (int)b; s++) pos_of_row[H.perm[s]] = s;
    });
    parallel_ranges(nb, [&](long long k0, long long k1, int) {
        for (int k = (int)k0; k < (int)k1; k++) {
            const int s0 = H.tile_ptr[k], s1 = H.tile_ptr[k + 1];
            BoxDesc &d = desc[k];
            d.nrows = s1 - s0;
            d.nlev = H.lev_off[k + 1] - H.lev_off[k] - 1;
            d.npred = H.pred_ptr[k + 1] - H.pred_ptr[k];
            d.next = 0;
            int w = 0;
            for (int s = s0; s < s1; s++) w = std::max(w, H.ptr[s + 1] - H.ptr[s]);
            for (int e = H.ptr[s0]; e < H.ptr[s1]; e++) d.next += (H.col[e] >= 0);
            d.nent = w;   // ELL width
            d.bytes = (int)ell_layout(d.nrows, w, d.nlev, d.next, d.npred).total;
        }
    }, 0, 64);
    size_t total = 0, cap = 0;
    int max_ext = 0;
    for (int k = 0; k < nb; k++) {
        desc[k].off = (long long)total;
        total += (size_t)desc[k].bytes;
        cap = std::max(cap, (size_t)desc[k].bytes);
        max_ext = std::max(max_ext, desc[k].next);
    }
    if (16 + cap + 8 * ((size_t)H.max_tile_rows + max_ext + 2) > (size_t)200 * 1024) return 2;
    P.blob.resize(std::max<size_t>(total, 16));
    if (total < 16) memset(P.blob.data(), 0, 16);
    unsigned char *blob = P.blob.data();
    const int np = host_threads();
    std::vector<char> bad(np, 0);
    parallel_ranges(nb, [&](long long k0, long long k1, int piece) {
        for (int k = (int)k0; k < (int)k1; k++) {
            const int s0 = H.tile_ptr[k];
            const BoxDesc &d = desc[k];
            const int w = d.nent, nr = d.nrows;
            const EllLayout lay = ell_layout(nr, w, d.nlev, d.next, d.npred);
            unsigned char *b = blob + d.off;
            memset(b, 0, (size_t)d.bytes);
            int *lev = (int *)(b + lay.lev), *perm = (int *)(b + lay.perm), *ext = (int *)(b + lay.ext), *pred = (int *)(b + lay.pred);
            int *ecol = (int *)(b + lay.ecol);
            double *diag = (double *)(b + lay.diag), *eval = (double *)(b + lay.eval);
            for (int L = 0; L <= d.nlev; L++) lev[L] = H.lev_ptr[H.lev_off[k] + L] - s0;
            // gate operand per predecessor box: of the rows this box reads from it, the one that box stores last
            for (int q = 0; q < d.npred; q++) {
                const int p = H.pred[H.pred_ptr[k] + q];
                int best = -1;
                for (int e = H.ptr[s0]; e < H.ptr[s0 + nr]; e++) {
                    const int c = H.col[e];
                    if (c < 0) continue;
                    const int pos = pos_of_row[c];
                    if (pos >= H.tile_ptr[p] && pos < H.tile_ptr[p + 1] && (best < 0 || pos > pos_of_row[best])) best = c;
                }
                if (best < 0) { bad[piece] = 1; return; }   // a predecessor without operands: the schedule is inconsistent
                pred[q] = best;
            }
            int q = 0;
            for (int s = 0; s < nr; s++) {
                perm[s] = H.perm[s0 + s];
                diag[s] = H.diag[s0 + s];
                const int e0 = H.ptr[s0 + s], len = H.ptr[s0 + s + 1] - e0;
                for (int j = 0; j < w; j++) {
                    if (j < len) {
                        const int c = H.col[e0 + j];
                        if (c >= 0) { ext[q] = c; ecol[j * nr + s] = nr + q; q++; }
                        else ecol[j * nr + s] = -c - 1;
                        eval[j * nr + s] = H.val[e0 + j];
                    }
                    else {
                        ecol[j * nr + s] = nr + d.next;   // the box's private +0.0 operand
                        eval[j * nr + s] = 0.0;
                    }
                }
            }
        }
    }, np, 1);
    for (char b : bad)
        if (b) return 1;
    P.desc_bytes.assign((const unsigned char *)desc.data(), (const unsigned char *)(desc.data() + desc.size()));
    P.cap = cap; P.max_ext = max_ext; P.flags = true;
    return 0;
}

// 0: packed; 2: not applicable (a box would not fit into shared memory, or the box graph is cyclic): slice schedule
int tri_tiled_pack_host(const TiledHost &H, PackedBoxes &P)
{
    if (!H.acyclic) return 2;
    return pack_ell_host(H, P);
}

int tri_tiled_upload(lsspg_ctx *ctx, const TiledHost &H, lsspg_tri *T)
{
    PackedBoxes P;
    const int rc = tri_tiled_pack_host(H, P);
    if (rc) return rc;
    return upload_common(ctx, H, T, P);
}

void tri_tiled_free(lsspg_tri *T)
{
    cudaFree(T->t_blob);
    cudaFree(T->t_desc);
    cudaFree(T->t_flags);
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

// Layout self-check for the CPU test-suite (never called by any product path): builds the
// tile schedule on the host and walks it box by box in ticket order, local level by local
// level, reading in-box operands from a private copy exactly as the kernel does.
// Returns 0 and *applicable = 1 when the factor has a tile schedule.
int lsspg_debug_tri_walk_tiled_host(int which, int n, const int *hTp, const int *hTj, const double *hTx, double *hx,
                                    const double *hrhs, int *applicable, int *info /* [8] */)
{
    TiledHost H;
    const int rc = tri_tiled_build_host(which, n, hTp, hTj, hTx, H);
    if (applicable) *applicable = (rc == 0);
    if (rc == 1) return 1;
    if (rc == 2) return 0;
    const double poison = strtod("nan", nullptr);
    for (int i = 0; i < n; i++) hx[i] = poison;   // an operand read before it is produced poisons the result
    // Boxes of one strongly connected component run concurrently on the device; emulate that by
    // sweeping over all unfinished boxes repeatedly: a box advances level by level, and inside
    // its current level every row whose operands have been produced is computed.
    std::vector<std::vector<double>> xs(H.num_tiles);
    std::vector<int> curL(H.num_tiles);
    std::vector<char> done(n, 0);
    int remaining = 0;
    for (int tk = 0; tk < H.num_tiles; tk++) {
        curL[tk] = H.lev_off[tk];
        xs[tk].assign(H.tile_ptr[tk + 1] - H.tile_ptr[tk], poison);
        if (curL[tk] < H.lev_off[tk + 1] - 1) remaining++;
    }
    while (remaining > 0) {
        bool progress = false;
        for (int tk = 0; tk < H.num_tiles; tk++) {
            const int s0 = H.tile_ptr[tk];
            while (curL[tk] < H.lev_off[tk + 1] - 1) {
                const int L = curL[tk];
                bool level_done = true;
                for (int slot = H.lev_ptr[L]; slot < H.lev_ptr[L + 1]; slot++) {
                    if (done[slot]) continue;
                    bool ready = true;
                    for (int e = H.ptr[slot]; e < H.ptr[slot + 1] && ready; e++) {
                        const int c = H.col[e];
                        const double v = c < 0 ? xs[tk][-c - 1] : hx[c];
                        if (v != v) ready = false;
                    }
                    if (!ready) { level_done = false; continue; }
                    double r = hrhs[H.perm[slot]];
                    for (int e = H.ptr[slot]; e < H.ptr[slot + 1]; e++) {
                        const int c = H.col[e];
                        r = r - H.val[e] * (c < 0 ? xs[tk][-c - 1] : hx[c]);
                    }
                    const double out = (H.diag[slot] == 1.0) ? r : r / H.diag[slot];
                    xs[tk][slot - s0] = out;
                    hx[H.perm[slot]] = out;
                    done[slot] = 1;
                    progress = true;
                }
                if (!level_done) break;
                curL[tk]++;
                if (curL[tk] == H.lev_off[tk + 1] - 1) remaining--;
            }
        }
        if (!progress) {
            lsspg::set_error("tiled walk: no progress (schedule would deadlock)");
            return 1;
        }
    }
    if (info) {
        info[0] = H.num_tiles; info[1] = H.num_tile_levels; info[2] = H.max_tile_rows; info[3] = H.num_levels;
        info[4] = H.grid_dims[0]; info[5] = H.grid_dims[1]; info[6] = H.grid_dims[2]; info[7] = H.tile_dims[0];
    }
    return 0;
}

// CPU emulation of the ELL box kernels FROM THE PACKED BLOBS (what the device reads), for the test-suite: boxes are
// visited round-robin in ticket order and a box advances by at most ONE chunk per round (the whole box when the blobs
// carry no chunk table), only when its gate rows and every operand of the chunk have been published.  x equals the serial
// sweep bit for bit when the image is right; info[0] = rounds needed = length of the longest chain of hand-offs,
// info[1] = chunks per box, info[2] = boxes, info[3] = levels of the box graph.  Never called by a product path.
int lsspg_debug_tri_walk_packed_host(int which, int n, const int *hTp, const int *hTj, const double *hTx, double *hx,
                                     const double *hrhs, int *applicable, int *info /* [4] */)
{
    TiledHost H;
    const int rc = tri_tiled_build_host(which, n, hTp, hTj, hTx, H);
    if (applicable) *applicable = 0;
    if (rc == 1) return 1;
    if (rc == 2) return 0;
    PackedBoxes P;
    {
        const int prc = tri_tiled_pack_host(H, P);
        if (prc == 1) return 1;
        if (prc == 2) return 0;
    }
    if (applicable) *applicable = 1;
    const BoxDesc *desc = (const BoxDesc *)P.desc_bytes.data();
    const double poison = strtod("nan", nullptr);
    for (int i = 0; i < n; i++) hx[i] = poison;
    const int nb = H.num_tiles;
    std::vector<std::vector<double>> sx(nb);
    std::vector<int> next_chunk(nb, 0);
    int remaining = nb, rounds = 0, max_chunks = 1;
    while (remaining > 0) {
        rounds++;
        int advanced = 0;
        std::vector<std::pair<int, double>> published;   // visible to the other boxes from the NEXT round on
        for (int tk = 0; tk < nb; tk++) {
            const BoxDesc &d = desc[tk];
            const unsigned char *b = P.blob.data() + d.off;
            const int nr = d.nrows, w = d.nent;
            const EllLayout lay = ell_layout(nr, w, d.nlev, d.next, d.npred);
            const int *lev = (const int *)(b + lay.lev), *perm = (const int *)(b + lay.perm), *ext = (const int *)(b + lay.ext);
            const int *tab = (const int *)(b + lay.pred), *ecol = (const int *)(b + lay.ecol);
            const double *diag = (const double *)(b + lay.diag), *eval = (const double *)(b + lay.eval);
            int C = 1;
            std::vector<int> cl = {0, d.nlev}, xp = {0, d.next}, gp = {0, d.npred};
            const int *gates = tab;
            max_chunks = std::max(max_chunks, C);
            const int c = next_chunk[tk];
            if (c >= C) continue;
            bool ready = true;
            for (int q = gp[c]; q < gp[c + 1] && ready; q++) ready = hx[gates[q]] == hx[gates[q]];
            for (int q = xp[c]; q < xp[c + 1] && ready; q++) ready = hx[ext[q]] == hx[ext[q]];
            if (!ready) continue;
            std::vector<double> &x = sx[tk];
            if (c == 0) {
                x.assign((size_t)nr + d.next + 2, poison);
                for (int s = 0; s < nr; s++) x[s] = hrhs[perm[s]];
                x[nr + d.next] = 0.0;
            }
            for (int q = xp[c]; q < xp[c + 1]; q++) x[nr + q] = hx[ext[q]];
            for (int L = cl[c]; L < cl[c + 1]; L++)
                for (int slot = lev[L]; slot < lev[L + 1]; slot++) {
                    double r = x[slot];
                    for (int k = 0; k < w; k++) r = r - eval[k * nr + slot] * x[ecol[k * nr + slot]];
                    if (diag[slot] != 1.0) r = r / diag[slot];
                    x[slot] = r;
                }
            for (int s = lev[cl[c]]; s < lev[cl[c + 1]]; s++) published.emplace_back(perm[s], x[s]);
            next_chunk[tk] = c + 1;
            advanced++;
            if (c + 1 == C) remaining--;
        }
        for (auto &pv : published) hx[pv.first] = pv.second;
        if (!advanced) {
            lsspg::set_error("packed walk: no progress (the schedule would deadlock)");
            return 1;
        }
    }
    if (info) { info[0] = rounds; info[1] = max_chunks; info[2] = nb; info[3] = H.num_tile_levels; }
    return 0;
}

}  // extern "C"
