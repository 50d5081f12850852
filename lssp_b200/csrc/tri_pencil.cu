// tri_pencil.cu -- pencil-streaming wavefront sweep for structured-grid triangular factors.
//
// Reference semantics: src/solver-tri.cxx:4-24 (lower, diagonal last) and :26-46 (upper, diagonal first, entries
// applied in DESCENDING storage order).  Every row subtracts its products one by one in the reference's order and
// divides by its stored diagonal, so the result is bit-identical; only the ORDER IN WHICH ROWS ARE VISITED changes.
//
// Why: an ILU(0) factor of an N^3 7-point grid has 3N-2 dependency levels.  Measured on B200
// (profiles/r02_ubench_latencies.txt): a dependent fp64 add or multiply takes 8 cycles, a shared-memory round trip
// 29, a 256-thread  LDS -> 3 subtractions -> STS -> bar.sync  step 117, but a hand-off between SMs through L2 ~550.
// The box schedule of round 1 (one warp per 8x8x8 box, whole-box hand-off) needed 9N in-box level steps of <= 48 rows
// and 3N/8 hops: 0.81 ms per sweep at 256^3 where the bytes would stream in 0.18 ms.
//
// Schedule.  The factor's offsets reveal a lattice (nx, ny, nz).  A LINE is the set of rows with fixed (y, z); a
// PENCIL is a tile of pv x pw lines (16 x 16 by default) and is owned by ONE CTA, one thread per line, which sweeps
// along x: at step k thread (v_l, w_l) computes the row x = k - v_l - w_l of its line -- a whole hyperplane of the
// pencil (up to 256 rows) per step, so that every dependency inside the pencil was computed at an earlier step and
// is read from a small ring of recent hyperplanes in shared memory.  Rows of OTHER pencils (the faces) arrive through
// MAILBOXES in global memory: the producing thread stores its value next to x, a helper warp of the consuming CTA
// polls the mailbox (the value is its own ready flag: mailboxes hold a sentinel NaN when empty), moves the value
// into the ring's ghost lanes and puts the sentinel back -- no fences, no flags, no pre-fill of x, and the consumer
// follows the producer at a distance of a few steps, not of a whole box.  Fill factors (ILU(k), k >= 1) have offsets
// with negative components; lines and pencils are then cut in the skewed coordinates of tri_tiled_build_host
// (u = x + s1 y + t1 z, v = y + s2 z, w = z), in which every dependency points backwards.
//
// HBM traffic.  A thread walks its line with unit stride, so neither column indices nor a permutation are read:
// the stream is the VALUES only, laid out [pencil][step][slot][thread] (coalesced, prefetched P steps ahead into
// registers), plus rhs (8 B) and x (8 B) per row: 40 B per row for the L of ILU(0) on a 7-point grid where the CSR
// sweep reads 68 B.  Missing entries (domain boundary) hold a NaN marker and multiply +0.0 by +0.0:
// r - (+0.0) == r bit for bit.
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <numeric>
#include "blas1.cuh"
#include "tri.cuh"
#include "host_par.h"

namespace lsspg {

constexpr unsigned long long kPenMissingBits = 0xFFF8C0DEFACE0002ull;   // "no such entry" in the value stream
constexpr unsigned long long kPenEmptyBits = 0xFFF8DEADBEEF0001ull;     // empty mailbox
constexpr int kPenG0 = -8;        // first virtual step the ghost prefetcher handles (ghost dk <= 8)
constexpr int kPenBatch = 4;      // virtual steps per prefetch round
constexpr int kPenNoOut = INT_MIN;
// shared-memory geometry of the kernel (see "the kernel" below)
constexpr int kPenChunk = 8;       // steps per rhs / x tile
constexpr int kPenNB = 3;          // rhs tiles
constexpr int kPenVRing = 8;       // steps of values in shared memory (even)
constexpr int kPenStagers = 128;   // threads
constexpr int kPenValuers = 64;    // threads
constexpr int kPenExtra = kPenStagers + kPenValuers + 32 + 32;   // + mailer + helper
constexpr int kPenL2Ahead = 32;    // steps the value stream is prefetched into L2

__host__ __device__ inline size_t pencil_smem_bytes(int T, int RS, int NV, int nb = kPenNB, int vr = kPenVRing)
{
    return sizeof(double) * ((size_t)kPenRing * RS + (size_t)nb * T * (kPenChunk + 1) + (size_t)vr * NV * T) +
           sizeof(int) * 2 * (size_t)T;
}


static inline double bits_to_double(unsigned long long b)
{
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// ---- host: schedule ------------------------------------------------------------------------------------------
int tri_pencil_build_host(int which, int n, const int *Tp, const int *Tj, const double *Tx, int num_sms, PencilHost &H)
{
    const bool lower = (which == LSSPG_TRI_LOWER);
    int g[3];
    std::vector<long long> offsets;
    if (n < 4096 || !detect_lattice(n, Tp, Tj, g, offsets)) return 2;
    const int W = (int)offsets.size();
    if (W < 1 || W > kPenMaxW) return 2;
    std::sort(offsets.begin(), offsets.end(), std::greater<long long>());   // slot order = application order (see below)
    // lattice vector of every offset: operand = row - (dx, dy, dz) in CANONICAL coordinates (upper factors: mirrored)
    long long dx[kPenMaxW], dy[kPenMaxW], dz[kPenMaxW];
    for (int w = 0; w < W; w++) {
        const long long d = offsets[w], plane = (long long)g[0] * g[1];
        long long z = d / plane, r = d % plane, y = r / g[0], x = r % g[0];
        if (x > g[0] / 2) { x -= g[0]; y += 1; }
        if (y > g[1] / 2 && g[2] > 1) { y -= g[1]; z += 1; }
        dx[w] = x; dy[w] = y; dz[w] = z;
    }
    // skew: smallest s1, s2, t1 >= 0 with  du = dx + s1 dy + t1 dz >= 0,  dv = dy + s2 dz >= 0  for every offset
    long long s1 = 0, s2 = 0, t1 = 0;
    {
        auto ceil_div = [](long long a, long long b) { return (a + b - 1) / b; };
        for (int w = 0; w < W; w++) {
            if (dz[w] == 0 && dy[w] > 0 && dx[w] < 0) s1 = std::max(s1, ceil_div(-dx[w], dy[w]));
            if (dz[w] > 0 && dy[w] < 0) s2 = std::max(s2, ceil_div(-dy[w], dz[w]));
        }
        for (int w = 0; w < W; w++)
            if (dz[w] > 0 && dx[w] + s1 * dy[w] < 0) t1 = std::max(t1, ceil_div(-(dx[w] + s1 * dy[w]), dz[w]));
        if (s1 > 4 || s2 > 4 || t1 > 8) return 2;
    }
    long long du[kPenMaxW], dv[kPenMaxW];
    for (int w = 0; w < W; w++) {
        du[w] = dx[w] + s1 * dy[w] + t1 * dz[w];
        dv[w] = dy[w] + s2 * dz[w];
        if (du[w] < 0 || dv[w] < 0 || dz[w] < 0 || du[w] + dv[w] + dz[w] < 1) return 2;
        if (du[w] + dv[w] + dz[w] > kPenMaxDk) return 2;
    }
    // every row: diagonal where the reference expects it, entries in application order = strictly increasing slot,
    // every entry a true lattice neighbour (no wrap-around across a grid line)
    const int np = host_threads();
    std::vector<char> bad(np, 0), nonunit(np, 0), hole(np, 0);
    auto canon = [&](int i, int &cx, int &cy, int &cz) {
        const int x = i % g[0], y = (i / g[0]) % g[1], z = i / (g[0] * g[1]);
        if (lower) { cx = x; cy = y; cz = z; }
        else { cx = g[0] - 1 - x; cy = g[1] - 1 - y; cz = g[2] - 1 - z; }
    };
    parallel_ranges(n, [&](long long r0, long long r1, int p) {
        for (int i = (int)r0; i < (int)r1; i++) {
            const int b = Tp[i], e = Tp[i + 1];
            if (e <= b) { bad[p] = 1; return; }
            const int dpos = lower ? e - 1 : b;
            if (Tj[dpos] != i) { bad[p] = 1; return; }
            if (Tx[dpos] != 1.0) nonunit[p] = 1;
            int cx, cy, cz, last = -1;
            canon(i, cx, cy, cz);
            auto inside = [&](int w) {
                return cx - dx[w] >= 0 && cx - dx[w] < g[0] && cy - dy[w] >= 0 && cy - dy[w] < g[1] && cz - dz[w] >= 0 &&
                       cz - dz[w] < g[2];
            };
            for (int q = 0; q < e - b - 1; q++) {
                const int k = lower ? b + q : e - 1 - q;   // src/solver-tri.cxx:17 / :39
                const long long d = lower ? (long long)i - Tj[k] : (long long)Tj[k] - i;
                int w = last + 1;
                while (w < W && offsets[w] != d) {
                    if (inside(w)) hole[p] = 1;   // the lattice neighbour exists but the row has no entry for it
                    w++;
                }
                if (d <= 0 || w >= W) { bad[p] = 1; return; }
                if (!inside(w)) { bad[p] = 1; return; }
                last = w;
            }
            for (int w = last + 1; w < W; w++)
                if (inside(w)) hole[p] = 1;
        }
    }, np);
    for (char b : bad)
        if (b) return 2;
    bool hasdiag = false, holes = false;
    for (char b : nonunit) hasdiag = hasdiag || b;
    // HOLES: some row lacks an entry although the neighbour exists (not the case for ILU(k) factors of uniform
    // stencils, where entries are only missing at the domain boundary).  Without holes a missing entry is stored as
    // +0.0 and its operand is +0.0 by construction (see the kernel): no test in the step loop.
    for (char b : hole) holes = holes || b;
    if (getenv("LSSPG_TRI_PENCIL_HOLES") && atoi(getenv("LSSPG_TRI_PENCIL_HOLES")) != 0) holes = true;
    // pencil cross-section
    int pv = 16, pw = 16;
    if (g[2] == 1) { pw = 1; pv = (g[1] >= 512) ? 64 : 32; }
    else {
        const int cand[4][2] = {{16, 16}, {16, 8}, {8, 8}, {8, 4}};
        for (int c = 0; c < 4; c++) {
            pv = cand[c][0]; pw = cand[c][1];
            const long long nvt = (g[1] + s2 * (g[2] - 1) + pv - 1) / pv, nwt = (g[2] + pw - 1) / pw;
            if (nvt * nwt >= num_sms) break;
        }
    }
    if (const char *e = getenv("LSSPG_TRI_PENCIL")) {
        int a, b;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0 && (a * b) % 32 == 0 && a * b <= kPenMaxThreads) { pv = a; pw = b; }
    }
    if (g[2] == 1) pw = 1;
    // the pencil must fit one SM: ring + rhs / x tiles + value ring (wide rows: fewer lines per pencil)
    while (pv * pw > 32 && pencil_smem_bytes(pv * pw, pencil_ring_stride(pv * pw, kPenMaxGhost), W + (hasdiag ? 1 : 0)) > (size_t)220 * 1024) {
        if (pw > 1 && pw >= pv / 2) pw /= 2;
        else pv /= 2;
    }
    const int T = pv * pw;
    if (T % 32 != 0 || T > kPenMaxThreads) return 2;
    const int NV = W + (hasdiag ? 1 : 0);
    // lines (canonical cy, cz) -> raw pencil, thread, first absolute step
    const int nlines = g[1] * g[2];
    const long long vext = g[1] + s2 * (g[2] - 1);
    const int nvt = (int)((vext + pv - 1) / pv), nwt = (g[2] + pw - 1) / pw;
    if ((long long)nvt * nwt > (1 << 24)) return 2;
    std::vector<int> line_raw(nlines), line_t(nlines), line_ks(nlines);
    for (int cz = 0; cz < g[2]; cz++)
        for (int cy = 0; cy < g[1]; cy++) {
            const long long v = cy + s2 * cz, w = cz;
            const int l = cz * g[1] + cy, vl = (int)(v % pv), wl = (int)(w % pw);
            line_raw[l] = (int)(w / pw) * nvt + (int)(v / pv);
            line_t[l] = vl + pv * wl;
            line_ks[l] = (int)(s1 * cy + t1 * cz) + vl + wl;
        }
    // non-empty pencils in ticket order: by diagonal of the pencil grid (every dependency points to a pencil with a
    // smaller V and/or Wt, so this order is topological and follows the wavefront)
    std::vector<int> raw_count((size_t)nvt * nwt, 0);
    for (int l = 0; l < nlines; l++) raw_count[line_raw[l]]++;
    std::vector<int> order;
    for (int r = 0; r < nvt * nwt; r++)
        if (raw_count[r]) order.push_back(r);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const int da = a / nvt + a % nvt, db = b / nvt + b % nvt;
        return da != db ? da < db : a < b;
    });
    const int npen = (int)order.size();
    std::vector<int> ticket_of_raw((size_t)nvt * nwt, -1);
    for (int k = 0; k < npen; k++) ticket_of_raw[order[k]] = k;
    std::vector<int> line_pen(nlines);
    std::vector<int> kmin(npen, INT_MAX), kmax(npen, INT_MIN);
    for (int l = 0; l < nlines; l++) {
        const int p = ticket_of_raw[line_raw[l]];
        line_pen[l] = p;
        kmin[p] = std::min(kmin[p], line_ks[l]);
        kmax[p] = std::max(kmax[p], line_ks[l] + g[0]);
    }
    H.hdr.assign(npen, PencilHdr());
    H.thr.assign((size_t)npen * T, PencilThread());
    long long val_total = 0;
    int max_steps = 0;
    for (int p = 0; p < npen; p++) {
        PencilHdr &h = H.hdr[p];
        h.nsteps = kmax[p] - kmin[p];
        h.thr_off = p * T;
        h.val_off = val_total;
        const long long padded = ((h.nsteps + 7) / 8) * 8 + 8;
        val_total += padded * NV * T;
        max_steps = std::max(max_steps, h.nsteps);
    }
    if (max_steps >= 32768) return 2;   // the kernel packs [kstart, kend) of a line into one word
    for (size_t q = 0; q < H.thr.size(); q++) {
        PencilThread &d = H.thr[q];
        d.kstart = 0; d.kend = 0; d.row0 = 0; d.pad = 0;
        for (int w = 0; w < kPenMaxW; w++) d.op[w] = (1 << 16) | kPenZeroLane;   // no such neighbour: the +0.0 lane
        for (int o = 0; o < kPenMaxOut; o++) d.out[o] = kPenNoOut;
    }
    // thread descriptors, ghost streams (consumer side)
    struct Stream { int pen, lane, src_line, sigma; long long mailbase; };
    std::vector<Stream> streams;
    H.ghost.clear();
    long long mail_len = 0;
    int max_ghost = 0, max_dk = 0;
    {
        std::vector<std::vector<int>> lines_of(npen);
        for (int l = 0; l < nlines; l++) lines_of[line_pen[l]].push_back(l);
        for (int p = 0; p < npen; p++) {
            PencilHdr &h = H.hdr[p];
            h.ghost_off = (int)H.ghost.size();
            std::map<int, int> lane_of_src;       // source line -> index into `mine`
            std::vector<Stream> mine;
            // pass 1: enumerate streams and their sigma = min(kstart + dx) - 1 over the consumers
            for (int l : lines_of[p]) {
                const int cy = l % g[1], cz = l / g[1];
                const int ks = line_ks[l] - kmin[p];
                for (int w = 0; w < W; w++) {
                    const long long oy = cy - dy[w], oz = cz - dz[w];
                    if (oy < 0 || oy >= g[1] || oz < 0 || oz >= g[2]) continue;
                    const int sl = (int)(oz * g[1] + oy);
                    if (line_pen[sl] == p) continue;
                    auto it = lane_of_src.find(sl);
                    if (it == lane_of_src.end()) {
                        lane_of_src[sl] = (int)mine.size();
                        mine.push_back({p, (int)mine.size(), sl, ks + (int)dx[w] - 1, 0});
                    }
                    else mine[it->second].sigma = std::min(mine[it->second].sigma, ks + (int)dx[w] - 1);
                }
            }
            if ((int)mine.size() > kPenMaxGhost) return 2;
            for (Stream &s : mine) {
                s.mailbase = mail_len;
                mail_len += g[0];
                PencilGhost gd;
                gd.mail0 = (int)(s.mailbase - s.sigma);
                gd.kg0 = s.sigma; gd.kg1 = s.sigma + g[0];
                gd.pad = 0;
                if (gd.kg0 < kPenG0) return 2;
                H.ghost.push_back(gd);
                streams.push_back(s);
            }
            if (mail_len > (long long)INT_MAX - (1 << 20)) return 2;
            h.nghost = (int)mine.size();
            h.kgend = h.nsteps;   // kgend: the prefetcher empties every mailbox position, also those no row reads
            for (int q = h.ghost_off; q < (int)H.ghost.size(); q++) h.kgend = std::max(h.kgend, H.ghost[q].kg1);
            max_ghost = std::max(max_ghost, h.nghost);
            // pass 2: descriptors
            for (int l : lines_of[p]) {
                const int cy = l % g[1], cz = l / g[1];
                const int ks = line_ks[l] - kmin[p];
                PencilThread &d = H.thr[(size_t)p * T + line_t[l]];
                d.kstart = ks; d.kend = ks + g[0];
                if (lower) d.row0 = l * g[0] - ks;
                else d.row0 = ((g[2] - 1 - cz) * g[1] + (g[1] - 1 - cy)) * g[0] + g[0] - 1 + ks;
                for (int w = 0; w < W; w++) {
                    const long long oy = cy - dy[w], oz = cz - dz[w];
                    if (oy < 0 || oy >= g[1] || oz < 0 || oz >= g[2]) continue;   // never an entry: slot stays a no-op
                    const int sl = (int)(oz * g[1] + oy);
                    int dk, lane;
                    if (line_pen[sl] == p) {
                        dk = line_ks[l] - line_ks[sl] + (int)dx[w];
                        lane = line_t[sl];
                    }
                    else {
                        const Stream &s = mine[lane_of_src[sl]];
                        dk = ks + (int)dx[w] - s.sigma;
                        lane = T + s.lane;
                    }
                    if (dk < 1 || dk > kPenMaxDk) return 2;
                    max_dk = std::max(max_dk, dk);
                    d.op[w] = (dk << 16) | lane;
                }
            }
        }
    }
    // producer side of every stream
    for (const Stream &s : streams) {
        const int sl = s.src_line, p = line_pen[sl];
        if (p >= s.pen) return 2;   // ticket order must be topological
        PencilThread &d = H.thr[(size_t)p * T + line_t[sl]];
        int o = 0;
        while (o < kPenMaxOut && d.out[o] != kPenNoOut) o++;
        if (o == kPenMaxOut) return 2;
        d.out[o] = (int)(s.mailbase - d.kstart);
    }
    // mailboxes per pencil (what the mailer warp walks), and whether the last slot is always the own line's predecessor
    H.outs.clear();
    bool own_last = true;
    for (int p = 0; p < npen; p++) {
        PencilHdr &h = H.hdr[p];
        h.out_off = (int)H.outs.size();
        for (int t = 0; t < T; t++) {
            const PencilThread &d = H.thr[(size_t)p * T + t];
            for (int o = 0; o < kPenMaxOut; o++)
                if (d.out[o] != kPenNoOut) H.outs.push_back({t, d.out[o], d.kstart, d.kend});
            if (d.kend > d.kstart && d.op[W - 1] != ((1 << 16) | t)) own_last = false;
        }
        h.nout = (int)H.outs.size() - h.out_off;
    }
    if (getenv("LSSPG_TRI_PENCIL_OWNLAST") && atoi(getenv("LSSPG_TRI_PENCIL_OWNLAST")) == 0) own_last = false;
    H.own_last = own_last;
    // the value stream
    H.vals.resize((size_t)val_total);
    {
        const double missing = holes ? bits_to_double(kPenMissingBits) : 0.0;
        double *vals = H.vals.data();
        parallel_ranges(val_total, [&](long long a, long long b, int) {
            for (long long q = a; q < b; q++) vals[q] = missing;
        });
        parallel_ranges(n, [&](long long r0, long long r1, int) {
            for (int i = (int)r0; i < (int)r1; i++) {
                int cx, cy, cz;
                canon(i, cx, cy, cz);
                const int l = cz * g[1] + cy, p = line_pen[l];
                const long long k = (long long)cx + line_ks[l] - kmin[p];
                double *base = vals + H.hdr[p].val_off + k * NV * T + line_t[l];
                const int b = Tp[i], e = Tp[i + 1];
                int w = 0;
                for (int q = 0; q < e - b - 1; q++) {
                    const int kk = lower ? b + q : e - 1 - q;
                    const long long d = lower ? (long long)i - Tj[kk] : (long long)Tj[kk] - i;
                    while (offsets[w] != d) w++;
                    base[(size_t)w * T] = Tx[kk];
                    w++;
                }
                if (hasdiag) base[(size_t)W * T] = lower ? Tx[e - 1] : Tx[b];
            }
        });
    }
    H.n = n; H.which = which; H.W = W; H.nv = NV; H.T = T; H.pv = pv; H.pw = pw; H.dir = lower ? 1 : -1;
    H.holes = holes;
    H.num_pencils = npen; H.max_ghost = max_ghost; H.max_dk = max_dk; H.max_steps = max_steps; H.mail_len = mail_len;
    H.offdiag_nnz = (long long)Tp[n] - n;
    for (int k = 0; k < 3; k++) H.grid_dims[k] = g[k];
    H.skew[0] = (int)s1; H.skew[1] = (int)s2; H.skew[2] = (int)t1;
    // number of dependency levels (reported by lsspg_tri_info): longest chain in steps
    {
        long long lev = 0;
        for (int w = 0; w < W; w++) (void)w;
        // rows (cx, cy, cz): level = max over lattice paths; for offsets with non-negative canonical components this is
        // the number of distinct hyperplanes; computed exactly by the slice builder when needed -- here an upper bound
        lev = (long long)g[0] + (s1 + 1) * (g[1] - 1) + (t1 + s2 + 1) * (g[2] - 1);
        H.num_levels = (int)std::min<long long>(lev, INT_MAX);
    }
    return 0;
}

// ---- host: replay of the packed image (CPU test-suite; never on a product path) ---------------------------------
// Pencils are visited in ticket order, each one completely (its producers have smaller tickets).  The ring, the
// ghost lanes, the mailboxes and the value stream are read exactly as the kernel reads them; an operand taken from
// an empty mailbox or from a ring slot that was overwritten poisons the result.
int tri_pencil_walk_host(const PencilHost &H, double *x, const double *rhs, int *info)
{
    const int T = H.T, W = H.W, NV = H.nv, RD = kPenRing, RS = pencil_ring_stride(T, H.max_ghost);
    const double empty = bits_to_double(kPenEmptyBits), poison = strtod("nan", nullptr);
    std::vector<double> mail((size_t)std::max<long long>(H.mail_len, 1), empty);
    std::vector<double> ring((size_t)RD * RS);
    std::vector<long long> ring_step((size_t)RD * RS);   // which step / virtual step the slot holds
    for (int i = 0; i < H.n; i++) x[i] = poison;
    long long bad_operands = 0;
    for (int p = 0; p < H.num_pencils; p++) {
        const PencilHdr &h = H.hdr[p];
        const PencilThread *thr = H.thr.data() + h.thr_off;
        const PencilGhost *gh = H.ghost.data() + h.ghost_off;
        std::fill(ring.begin(), ring.end(), 0.0);          // the kernel zeroes the ring of every pencil
        std::fill(ring_step.begin(), ring_step.end(), LLONG_MIN);
        std::vector<double> out(T);
        int ghost_ready = kPenG0;
        for (int k = 0; k < h.nsteps; k++) {
            // the prefetcher is at least up to virtual step k - 1 and at most kPenRing - kPenMaxDk - 1 ahead; replay the
            // furthest-ahead case, which is the one that can overwrite slots still in use
            while (ghost_ready < ((k == h.nsteps - 1) ? h.kgend : std::min(k + RD - kPenMaxDk, h.kgend))) {
                const int kg = ghost_ready;
                for (int e = 0; e < h.nghost; e++) {
                    const size_t slot = (size_t)(kg & (RD - 1)) * RS + T + e;
                    if (kg < gh[e].kg0 || kg >= gh[e].kg1) { ring[slot] = 0.0; ring_step[slot] = LLONG_MIN; continue; }
                    double &m = mail[(size_t)gh[e].mail0 + kg];
                    unsigned long long mb;
                    memcpy(&mb, &m, 8);
                    if (mb == kPenEmptyBits) { set_error("pencil walk: mailbox empty (ticket order is not topological)"); return 1; }
                    ring[slot] = m; ring_step[slot] = kg;
                    m = empty;
                }
                ghost_ready++;
            }
            for (int t = 0; t < T; t++) {
                const PencilThread &d = thr[t];
                const bool act = k >= d.kstart && k < d.kend;
                double r = 0.0;
                if (act) {
                    const double *v = H.vals.data() + h.val_off + (size_t)k * NV * T + t;
                    r = rhs[d.row0 + H.dir * k];
                    for (int w = 0; w < W; w++) {
                        const double a = v[(size_t)w * T];
                        unsigned long long ab;
                        memcpy(&ab, &a, 8);
                        const int lane = (d.op[w] & 0xffff) == kPenZeroLane ? RS - 1 : (d.op[w] & 0xffff), dk = d.op[w] >> 16;
                        const size_t slot = (size_t)((k - dk) & (RD - 1)) * RS + lane;
                        double pr;
                        if (H.holes && ab == kPenMissingBits) pr = 0.0;
                        else {
                            // a stored +0.0 in a hole-free factor is a missing entry: its operand must be +0.0 exactly
                            if (!H.holes && ab == 0 && lane != RS - 1 && ring_step[slot] != k - dk) {
                                unsigned long long xb;
                                memcpy(&xb, &ring[slot], 8);
                                if (xb != 0) bad_operands++;
                            }
                            else if (!(ab == 0 && !H.holes) && ring_step[slot] != k - dk) bad_operands++;
                            pr = a * ring[slot];
                        }
                        r = r - pr;   // src/solver-tri.cxx:18 / :40
                    }
                    if (NV > W) r = r / v[(size_t)W * T];   // :22 / :44
                }
                out[t] = act ? r : 0.0;
            }
            for (int t = 0; t < T; t++) {
                const PencilThread &d = thr[t];
                const size_t slot = (size_t)(k & (RD - 1)) * RS + t;
                ring[slot] = out[t];
                ring_step[slot] = (k >= d.kstart && k < d.kend) ? k : LLONG_MIN;
                if (k >= d.kstart && k < d.kend) {
                    x[d.row0 + H.dir * k] = out[t];
                    for (int o = 0; o < kPenMaxOut; o++)
                        if (d.out[o] != kPenNoOut) mail[(size_t)d.out[o] + k] = out[t];
                }
            }
        }
    }
    if (bad_operands) { set_error("pencil walk: %lld operands read from a ring slot that does not hold them", bad_operands); return 1; }
    // every mailbox must be empty again (the next application relies on it)
    for (long long q = 0; q < H.mail_len; q++) {
        unsigned long long mb;
        memcpy(&mb, &mail[(size_t)q], 8);
        if (mb != kPenEmptyBits) { set_error("pencil walk: mailbox %lld not emptied", q); return 1; }
    }
    if (info) { info[0] = H.num_pencils; info[1] = H.T; info[2] = H.max_ghost; info[3] = H.max_dk; info[4] = H.W; info[5] = H.nv; info[6] = H.max_steps; info[7] = H.skew[0] * 100 + H.skew[1] * 10 + H.skew[2] + (H.holes ? 1000 : 0); }
    return 0;
}

// ---- device ----------------------------------------------------------------------------------------------------
struct PencilArgs {
    const PencilHdr *hdr;
    const PencilThread *thr;
    const PencilGhost *ghost;
    const PencilOut *outs;
    const double *vals;
    double *mail;
    unsigned int *counter;
    int num_pencils, T, RS, dir;
    int nb, vr, l2ahead;        // rhs tiles and steps of values kept in shared memory; steps the value stream is prefetched into L2
    double *x;
    const double *rhs;
    const int *stop;
    int *err;
    int dbg;                    // LSSPG_TRI_PENCIL_DBG: timing experiments only (skips work, wrong results)
    unsigned long long *prof;   // LSSPG_TRI_PROF=1: 8 words per pencil (start ns, end ns, cycles: total, ghost wait, barrier; steps)
};

__device__ __forceinline__ unsigned long long pen_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double pen_ld_relaxed(const double *p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pen_st_relaxed(double *p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// value stream: read once, keep it out of L1
__device__ __forceinline__ double pen_ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// Progress words in shared memory, written by one role and polled by another (compute warps, stagers, valuers, mailer,
// helper).  They are release / acquire at CTA scope: the data they announce (ring rows, tiles, value ring) is written by
// OTHER threads of the announcing role, ordered before the announcement by that role's barrier -- release is cumulative
// over it.  (An earlier version used plain volatile accesses: intermittently wrong sweeps at 256^3, 50 000 of 16.8 M rows,
// scripts/dist_pc_check.py.  The very first version's acquire loads were slow only because the compute threads then had
// global loads in flight, which the fence drained; they have none any more.)
__device__ __forceinline__ int pen_ld_flag(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned int)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void pen_st_flag(int *p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// One step of a pencil is 256 rows on 256 DIFFERENT grid lines (2 KB apart).  What was measured on the way here
// (256^3, L sweep; profiles/r02_pencil_*.txt, scripts/ubench/step.cu):
//   * per-thread rhs loads / x stores touch 32 cache lines per warp request (~66 cycles of L1 replays): 0.95 ms;
//   * rhs / x staged through shared memory by the compute threads themselves: a dependent stream of ~100 instructions
//     per warp and step, values fetched into registers aliasing on the warp's 6 scoreboards: 0.59 ms;
//   * a step is bound by SHARED-MEMORY BANDWIDTH AND LATENCY: every 64-bit LDS / STS of the 8 compute warps costs ~16
//     cycles (128 B/clk), an exposed poll of a progress word ~80, a mailbox store by two lanes of a warp ~40.
// Hence WARP SPECIALISATION with as few shared-memory accesses as possible on the compute warps, which keep only
//     LDS operands -> products -> subtractions (-> divide) -> STS -> bar.sync
// while everything that moves data runs beside them, synchronised through release / acquire progress words in shared
// memory (polled one step ahead of their use, so that no poll latency is exposed):
//   * stagers (4 warps): rhs tiles in ([line][8 + 1] doubles, cp.async, consecutive threads along x = 64-byte pieces of
//     a line); x out, read straight from the ring of hyperplanes;
//   * valuers (2 warps): the value stream [step][slot][thread] into a ring of kPenVRing steps, 16-byte cp.async;
//   * mailer (1 warp): face values from the ring to the mailboxes of the pencils that read them;
//   * helper (1 warp): mailboxes -> ghost lanes, and prefetch.global.L2 of the value stream kPenL2Ahead steps ahead.
enum { CTL_TICKET = 0, CTL_GHOST = 1, CTL_STEPS = 2, CTL_ABORT = 3, CTL_RHS = 4, CTL_VALS = 5, CTL_MAILED = 6, CTL_XOUT = 7 };

// spin until *flag >= want (progress words only grow); a raised abort flag ends every wait so that all roles run to
// the end of the pencil with matching barrier counts (the sweep is then wrong and FLAG_TRI_TIMEOUT says so)
__device__ __forceinline__ void pen_wait(const int *flag, int want, const int *abort_flag)
{
    while (pen_ld_flag(flag) < want)
        if (pen_ld_flag(abort_flag)) break;
}
// the same for the data-moving roles: they sleep between polls so that their spinning neither takes issue slots from
// the compute warps on the same scheduler nor fills the shared-memory pipe with polls
__device__ __forceinline__ void pen_wait_sleep(const int *flag, int want, const int *abort_flag)
{
    while (pen_ld_flag(flag) < want) {
        if (pen_ld_flag(abort_flag)) break;
        __nanosleep(64);
    }
}

template <int W, int DIAG, bool HOLES, bool OWN, bool PROF>
__global__ void __launch_bounds__(kPenMaxThreads + kPenExtra, 2) tri_pencil_kernel(const PencilArgs a)
{
    extern __shared__ __align__(16) double smem_d[];
    __shared__ int s_ctl[8];
    constexpr int NV = W + DIAG, RD = kPenRing, C = kPenChunk, CP = kPenChunk + 1;
    const int VR = a.vr, NB = a.nb;
    constexpr int WR = OWN ? W - 1 : W;               // operands read from the ring
    const int T = a.T, RS = a.RS, tid = threadIdx.x;
    double *vring = smem_d;                           // [VR][NV][T] (first: 16-byte aligned for cp.async)
    double *rhsT = vring + (size_t)VR * NV * T;       // [kPenNB][T][C + 1]
    double *ring = rhsT + (size_t)NB * T * CP;        // [RD][RS]: lanes [0, T) own lines, then ghost lines, last: +0.0
    int2 *s_line = reinterpret_cast<int2 *>(ring + (size_t)RD * RS);   // per line: row0, kstart | length << 16
    if (a.stop && *a.stop) return;
    if (tid == 0) s_ctl[CTL_ABORT] = 0;
    const unsigned int total = (unsigned int)a.num_pencils + gridDim.x;
    const long long dir = a.dir;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_ctl[CTL_TICKET] = (int)atomicInc(a.counter, total - 1);
            s_ctl[CTL_GHOST] = kPenG0;
            s_ctl[CTL_STEPS] = 0;
            s_ctl[CTL_RHS] = 0;
            s_ctl[CTL_VALS] = 0;
            s_ctl[CTL_MAILED] = 0;
            s_ctl[CTL_XOUT] = 0;
        }
        // Operands of rows that do not exist (line starts, domain boundary, virtual steps a ghost line does not have)
        // are read from ring slots nobody writes: they must hold +0.0 so that the +0.0 stored for the missing entry
        // gives the product +0.0 and r - (+0.0) == r bit for bit.
        for (int q = tid; q < RD * RS; q += blockDim.x) ring[q] = 0.0;
        __syncthreads();
        const unsigned int tk = (unsigned int)s_ctl[CTL_TICKET];
        if (tk >= (unsigned int)a.num_pencils || s_ctl[CTL_ABORT]) break;
        const PencilHdr h = a.hdr[tk];
        const int nchunks = (h.nsteps + C - 1) / C;
        int4 d0 = make_int4(0, 0, 0, 0), d1 = d0, d2 = d0;
        if (tid < T) {
            const int4 *dp = reinterpret_cast<const int4 *>(a.thr + h.thr_off + tid);
            d0 = __ldg(dp); d1 = __ldg(dp + 1); d2 = __ldg(dp + 2);
            s_line[tid] = make_int2(d0.z, d0.x | ((d0.y - d0.x) << 16));
        }
        __syncthreads();   // the line table is there for the stagers
        if (tid < T) {
            // ---- compute threads: one line each ----
            const int kstart = d0.x, klen = d0.y - d0.x;
            const int opw[8] = {d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w};
            const unsigned int RS8 = (unsigned int)RS * 8u, ring_bytes = (unsigned int)RD * RS8;
            char *ringb = reinterpret_cast<char *>(ring);
            unsigned int rop[WR > 0 ? WR : 1];   // byte offset in the ring of slot w's operand for the current step
#pragma unroll
            for (int w = 0; w < WR; w++) {
                const int lane = ((opw[w] & 0xffff) == kPenZeroLane) ? RS - 1 : (opw[w] & 0xffff);
                rop[w] = (unsigned int)((0 - (opw[w] >> 16)) & (RD - 1)) * RS8 + (unsigned int)lane * 8u;
            }
            unsigned int rown = (unsigned int)tid * 8u;   // own slot of the current step
            const unsigned int vstep = (unsigned int)NV * T * 8u, vring_bytes = (unsigned int)VR * vstep;
            const char *vb = reinterpret_cast<const char *>(vring + tid);
            unsigned int voff = 0;                          // byte offset of the current step in the value ring
            unsigned long long p_t0 = 0, p_c0 = 0, p_wait = 0, p_bar = 0, p_wv = 0, p_wr = 0;
            const bool prof = PROF && a.prof != nullptr && tid == 0;
            if (prof) { p_t0 = pen_globaltimer(); p_c0 = clock64(); }
            // Progress words are read one poll period BEFORE they are needed (they only grow, so a stale value that
            // already suffices is good); only a value that does not suffice is polled again, in a loop.
            int f_vals = 0, f_ghost = kPenG0, f_rhs = 0, f_mail = 0, f_xout = 0;
            double prev = 0.0;                              // OWN: the row's predecessor on its own line
            int k = 0, buf = 0;
            for (int c = 0; c < nchunks; c++) {
                {   // chunk c needs its rhs tile; its ring rows are those of chunk c - 2: mailed and written out
                    unsigned long long t = 0;
                    if (prof) t = clock64();
                    if (f_rhs < c + 1) pen_wait(&s_ctl[CTL_RHS], c + 1, &s_ctl[CTL_ABORT]);
                    if (f_xout < c - 1) pen_wait(&s_ctl[CTL_XOUT], c - 1, &s_ctl[CTL_ABORT]);
                    if (h.nout && f_mail < (c - 1) * C) pen_wait(&s_ctl[CTL_MAILED], (c - 1) * C, &s_ctl[CTL_ABORT]);
                    if (prof) p_wr += clock64() - t;
                }
                const double *rt = rhsT + ((size_t)buf * T + tid) * CP;
#pragma unroll
                for (int j = 0; j < C; j++) {
                    if ((j & 1) == 0) {   // values of steps k, k + 1
                        const int want = min(k + 2, h.nsteps);
                        if (f_vals < want) {
                            unsigned long long t = 0;
                            if (prof) t = clock64();
                            pen_wait(&s_ctl[CTL_VALS], want, &s_ctl[CTL_ABORT]);
                            if (prof) p_wv += clock64() - t;
                        }
                        f_vals = pen_ld_flag(&s_ctl[CTL_VALS]);   // consumed two steps from now
                    }
                    if ((j & (kPenBatch - 1)) == 0 && h.nghost) {   // ghost lanes of steps k .. k + 3: virtual steps < k + 4
                        const int want = min(k + kPenBatch, h.nsteps);
                        if (f_ghost < want) {
                            unsigned long long t = 0;
                            if (prof) t = clock64();
                            pen_wait(&s_ctl[CTL_GHOST], want, &s_ctl[CTL_ABORT]);
                            if (prof) p_wait += clock64() - t;
                        }
                        f_ghost = pen_ld_flag(&s_ctl[CTL_GHOST]);
                    }
                    if (j == C - 2) { f_rhs = pen_ld_flag(&s_ctl[CTL_RHS]); f_mail = pen_ld_flag(&s_ctl[CTL_MAILED]); f_xout = pen_ld_flag(&s_ctl[CTL_XOUT]); }
                    double acc = rt[j];
                    double av[NV];
#pragma unroll
                    for (int w = 0; w < NV; w++) av[w] = *reinterpret_cast<const double *>(vb + voff + (unsigned int)w * (unsigned int)T * 8u);
                    voff += vstep;
                    if (voff >= vring_bytes) voff = 0;
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        double xv;
                        if (OWN && w == W - 1) xv = prev;
                        else {
                            xv = *reinterpret_cast<const double *>(ringb + rop[w < WR ? w : 0]);
                            rop[w < WR ? w : 0] += RS8;
                            if (rop[w < WR ? w : 0] >= ring_bytes) rop[w < WR ? w : 0] -= ring_bytes;
                        }
                        double pr = av[w] * xv;
                        if (HOLES && __double2hiint(av[w]) == (int)(kPenMissingBits >> 32)) pr = 0.0;   // r - (+0.0) == r
                        acc = acc - pr;                    // src/solver-tri.cxx:18 / :40
                    }
                    if (DIAG) acc = acc / av[NV - 1];      // :22 / :44
                    const bool act = (unsigned int)(k - kstart) < (unsigned int)klen;
                    const double out = act ? acc : 0.0;
                    prev = out;
                    *reinterpret_cast<double *>(ringb + rown) = out;
                    rown += RS8;
                    if (rown >= ring_bytes) rown -= ring_bytes;
                    k++;
                    {   // (one bar.sync for the whole warp: the aligned barrier must not be executed divergently)
                        unsigned long long t = 0;
                        if (PROF && prof) t = clock64();
                        asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
                        if (PROF && prof) p_bar += clock64() - t;
                    }
                    if (tid == 0) pen_st_flag(&s_ctl[CTL_STEPS], k);
                }
                buf = (buf + 1 >= NB) ? 0 : buf + 1;
            }
            if (prof) {
                unsigned long long *q = a.prof + 8 * (size_t)tk;
                q[0] = p_t0; q[1] = pen_globaltimer(); q[2] = clock64() - p_c0; q[3] = p_wait; q[4] = p_bar; q[5] = (unsigned long long)h.nsteps;
                q[6] = p_wv; q[7] = p_wr;
            }
        }
        else if (tid < T + kPenStagers) {
            // ---- stagers: rhs tiles in (cp.async), x out of the ring.  Element e of a chunk: line e / 8, column e % 8 ----
            const int sid = tid - T, per = T * C / kPenStagers;
            auto rhs_in = [&](int chunk, int buf) {
                const unsigned int tile = (unsigned int)__cvta_generic_to_shared(rhsT + (size_t)buf * T * CP);
#pragma unroll 4
                for (int i = 0; i < per; i++) {
                    const int e = i * kPenStagers + sid, line = e >> 3, col = e & 7, kk = chunk * C + col;
                    const int2 ln = s_line[line];
                    const bool act = (unsigned int)(kk - (ln.y & 0xffff)) < (unsigned int)(ln.y >> 16);
                    const double *src = act ? a.rhs + ((long long)ln.x + dir * kk) : a.rhs;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(tile + (unsigned int)((line * CP + col) * 8)), "l"(src), "r"(act ? 8 : 0) : "memory");
                }
            };
            for (int c0 = 0; c0 < NB && c0 < nchunks; c0++) rhs_in(c0, c0);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 2, %0;" ::"n"(kPenStagers) : "memory");
            if (sid == 0) pen_st_flag(&s_ctl[CTL_RHS], NB);
            for (int i = 1; i <= nchunks; i++) {
                // Chunk i - 1 has been computed.  (1) The tiles requested so far (up to chunk i + NB - 2) have landed: publish.
                // (2) The tile chunk i - 1 used is free: request chunk i + NB - 1 into it.  (3) x of chunk i - 1 goes out of
                // the ring (rows (i - 1) C .. + 7, which chunk i + 1 will reuse: the compute threads wait for CTL_XOUT >= i).
                pen_wait_sleep(&s_ctl[CTL_STEPS], i * C, &s_ctl[CTL_ABORT]);
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                asm volatile("bar.sync 2, %0;" ::"n"(kPenStagers) : "memory");
                if (sid == 0) pen_st_flag(&s_ctl[CTL_RHS], i + NB - 1);
                if (i + NB - 1 < nchunks) rhs_in(i + NB - 1, (i - 1) % NB);
                asm volatile("cp.async.commit_group;" ::: "memory");
                for (int q0 = 0; q0 < per; q0 += 8) {   // 8 elements in flight per thread
                    int2 ln[8];
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) ln[u] = s_line[min((q0 + u) * kPenStagers + sid, T * C - 1) >> 3];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int e = min((q0 + u) * kPenStagers + sid, T * C - 1), kk = (i - 1) * C + (e & 7);
                        v[u] = ring[(kk & (RD - 1)) * RS + (e >> 3)];
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int e = (q0 + u) * kPenStagers + sid, kk = (i - 1) * C + (e & 7);
                        if (q0 + u < per && (unsigned int)(kk - (ln[u].y & 0xffff)) < (unsigned int)(ln[u].y >> 16))
                            a.x[(long long)ln[u].x + dir * kk] = v[u];
                    }
                }
                asm volatile("bar.sync 2, %0;" ::"n"(kPenStagers) : "memory");
                if (sid == 0) pen_st_flag(&s_ctl[CTL_XOUT], i);
            }
        }
        else if (tid < T + kPenStagers + kPenValuers) {
            // ---- valuers: the value stream into the ring, two steps (one contiguous piece) per round ----
            const int vid = tid - T - kPenStagers;
            const int npairs = (h.nsteps + 1) / 2;
            const int pieces = NV * T;                       // 16-byte pieces per pair
            const char *src0 = reinterpret_cast<const char *>(a.vals + h.val_off);
            const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(vring);
            const unsigned int pair_bytes = 2u * NV * T * 8u;
            // pair q is published `lag` rounds after it was requested; lag < VR / 2, or the request of pair q + lag (which
            // waits for the slot of pair q + lag - VR / 2 to be consumed) would stand before the publication it waits for
            const int lag = (VR >= 8) ? 2 : 1;
            for (int q = 0; q < npairs + lag; q++) {
                if (q < npairs) {
                    pen_wait_sleep(&s_ctl[CTL_STEPS], 2 * q - (VR - 2), &s_ctl[CTL_ABORT]);   // the ring slot of pair q is free
                    const char *src = src0 + (size_t)q * pair_bytes;
                    const unsigned int dst = ring_s + (unsigned int)(q % (VR / 2)) * pair_bytes;
                    for (int i = vid; i < pieces; i += kPenValuers) {
                        const unsigned int o = (unsigned int)i * 16u;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (lag == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");   // pair q - lag has landed (this thread's pieces)
                else asm volatile("cp.async.wait_group 1;" ::: "memory");
                asm volatile("bar.sync 3, %0;" ::"n"(kPenValuers) : "memory");      // ... and everybody's
                if (vid == 0 && q >= lag) pen_st_flag(&s_ctl[CTL_VALS], 2 * (q - lag) + 2);
            }
        }
        else if (tid < T + kPenStagers + kPenValuers + 32) {
            // ---- mailer: the face values of every finished step go to the mailboxes of the pencils that read them ----
            const int lane = tid - T - kPenStagers - kPenValuers;
            const PencilOut *po = a.outs + h.out_off;
            int done = 0;
            while (done < h.nsteps && h.nout) {
                int sd;
                while ((sd = pen_ld_flag(&s_ctl[CTL_STEPS])) <= done) {
                    if (pen_ld_flag(&s_ctl[CTL_ABORT])) { sd = h.nsteps; break; }
                    __nanosleep(32);
                }
                sd = min(sd, h.nsteps);
                for (int e = lane; e < h.nout; e += 32) {
                    const PencilOut o = po[e];
                    for (int kk = max(done, o.kstart); kk < min(sd, o.kend); kk++)
                        pen_st_relaxed(a.mail + ((long long)o.mail0 + kk), ring[(kk & (RD - 1)) * RS + o.lane]);
                }
                done = sd;
                __syncwarp();
                if (lane == 0) pen_st_flag(&s_ctl[CTL_MAILED], done);
            }
        }
        else {
            // ---- helper warp: value stream -> L2, mailboxes -> ghost lanes; kPenBatch (virtual) steps per round ----
            const int lane = tid - T - kPenStagers - kPenValuers - 32;
            const PencilGhost *gh = a.ghost + h.ghost_off;
            bool aborted = false;
            const char *vbase = reinterpret_cast<const char *>(a.vals + h.val_off);
            const long long step_bytes = (long long)NV * T * 8, vend = step_bytes * h.nsteps;
            for (long long b = (long long)lane * 128; b < step_bytes * a.l2ahead && b < vend; b += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vbase + b));
            for (int g = kPenG0; g < h.kgend; g += kPenBatch) {
                // a ghost slot is reused every RD virtual steps and read up to kPenMaxDk steps after it was written
                {
                    int sd;
                    while ((sd = pen_ld_flag(&s_ctl[CTL_STEPS])) < h.nsteps && sd + (RD - kPenMaxDk) < g + kPenBatch) {
                        if (pen_ld_flag(&s_ctl[CTL_ABORT])) break;
                        __nanosleep(64);
                    }
                }
                if (g >= 0) {   // the values of steps [g + kPenL2Ahead, + kPenBatch) on their way to L2
                    const long long b0 = step_bytes * (g + a.l2ahead), b1 = min(b0 + step_bytes * kPenBatch, vend);
                    for (long long b = b0 + (long long)lane * 128; b < b1; b += 32 * 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(vbase + b));
                }
                for (int e = lane; e < h.nghost; e += 32) {
                    const PencilGhost gd = gh[e];
                    double val[kPenBatch];
                    bool need[kPenBatch];
#pragma unroll
                    for (int j = 0; j < kPenBatch; j++) {
                        need[j] = (g + j >= gd.kg0) && (g + j < gd.kg1) && !aborted;
                        val[j] = need[j] ? pen_ld_relaxed(a.mail + ((long long)gd.mail0 + g + j)) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < kPenBatch; j++) {
                        int spins = 0;
                        while (need[j] && !aborted && (unsigned long long)__double_as_longlong(val[j]) == kPenEmptyBits) {
                            if (++spins > 8) __nanosleep(spins > 64 ? 200 : 40);
                            if (spins > (1 << 21) || pen_ld_flag(&s_ctl[CTL_ABORT])) { aborted = true; break; }
                            val[j] = pen_ld_relaxed(a.mail + ((long long)gd.mail0 + g + j));
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kPenBatch; j++) {
                        ring[((g + j) & (RD - 1)) * RS + T + e] = val[j];   // +0.0 where the ghost line has no row
                        if (need[j] && !aborted)
                            pen_st_relaxed(a.mail + ((long long)gd.mail0 + g + j), __longlong_as_double((long long)kPenEmptyBits));
                    }
                }
                if (__any_sync(0xffffffffu, aborted) && !pen_ld_flag(&s_ctl[CTL_ABORT])) {
                    aborted = true;
                    if (lane == 0) { *a.err = 1; pen_st_flag(&s_ctl[CTL_ABORT], 1); }
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) pen_st_flag(&s_ctl[CTL_GHOST], g + kPenBatch);
            }
        }
    }
}

template <int W, int DIAG, bool HOLES, bool OWN, bool PROF>
static int pencil_launch(lsspg_ctx *ctx, const lsspg_tri *Tr, const PencilArgs &a_in)
{
    auto kern = tri_pencil_kernel<W, DIAG, HOLES, OWN, PROF>;
    PencilArgs a = a_in;
    // LSSPG_TRI_PENCIL_LEAN=1: two pencils per SM with shallower buffers (2 rhs tiles, 4 steps of values), so that pencils
    // waiting for their predecessors share an SM with pencils that work.  Measured at 256^3 (all 256 pencils resident):
    // SLOWER, 0.69 / 0.84 ms instead of 0.49 / 0.55 -- the shallow value ring stalls every step; off by default.
    static int env_lean = -1;
    if (env_lean < 0) env_lean = getenv("LSSPG_TRI_PENCIL_LEAN") ? atoi(getenv("LSSPG_TRI_PENCIL_LEAN")) : 0;
    a.nb = kPenNB; a.vr = kPenVRing; a.l2ahead = kPenL2Ahead;
    {   // tuning knobs (even VR >= 4, NB >= 2); defaults are the measured optimum at 256^3
        static int env_vr = -1, env_nb = -1, env_l2 = -1;
        if (env_vr < 0) {
            env_vr = getenv("LSSPG_TRI_PENCIL_VR") ? atoi(getenv("LSSPG_TRI_PENCIL_VR")) : 0;
            env_nb = getenv("LSSPG_TRI_PENCIL_NB") ? atoi(getenv("LSSPG_TRI_PENCIL_NB")) : 0;
            env_l2 = getenv("LSSPG_TRI_PENCIL_L2") ? atoi(getenv("LSSPG_TRI_PENCIL_L2")) : 0;
        }
        if (env_vr >= 4 && env_vr % 2 == 0 && env_vr <= 64) a.vr = env_vr;
        if (env_nb >= 2 && env_nb <= 8) a.nb = env_nb;
        if (env_l2 > 0) a.l2ahead = env_l2;
        while (a.vr > 4 && pencil_smem_bytes(a.T, a.RS, W + DIAG, a.nb, a.vr) > (size_t)226 * 1024) a.vr -= 2;
    }
    if (env_lean && pencil_smem_bytes(a.T, a.RS, W + DIAG, 2, 4) <= (size_t)112 * 1024 &&
        pencil_smem_bytes(a.T, a.RS, W + DIAG) > (size_t)112 * 1024 && Tr->num_tiles > ctx->num_sms) { a.nb = 2; a.vr = 4; }
    const size_t smem = pencil_smem_bytes(a.T, a.RS, W + DIAG, a.nb, a.vr);
    const int block = a.T + kPenExtra;
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        LSSPG_CHECK(smem <= (size_t)226 * 1024, "tri_pencil: %zu bytes of shared memory per pencil exceed one SM", smem);
        LSSPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int occ = 0;
    LSSPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    LSSPG_CHECK(occ >= 1, "tri_pencil: the kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    static int env_cap = -1;
    if (env_cap < 0) {
        const char *e = getenv("LSSPG_TRI_PENCIL_CTAS_PER_SM");
        env_cap = e ? std::max(1, atoi(e)) : 8;
    }
    occ = std::min(occ, env_cap);
    const int grid = std::max(1, std::min(Tr->num_tiles, ctx->num_sms * occ));
    LSSPG_LAUNCH(ctx, kern, grid, block, smem, a);
    return 0;
}

template <int W, int DIAG>
static int pencil_dispatch2(lsspg_ctx *ctx, const lsspg_tri *Tr, const PencilArgs &a)
{
    if (Tr->p_holes) return pencil_launch<W, DIAG, true, false, false>(ctx, Tr, a);
    if (!Tr->p_own_last) return pencil_launch<W, DIAG, false, false, false>(ctx, Tr, a);
    if (a.prof && W == 3) return pencil_launch<W, DIAG, false, true, (W == 3)>(ctx, Tr, a);   // profiling build: 7-point ILU(0) shapes only
    return pencil_launch<W, DIAG, false, true, false>(ctx, Tr, a);
}

template <int W>
static int pencil_dispatch(lsspg_ctx *ctx, const lsspg_tri *Tr, const PencilArgs &a)
{
    return Tr->p_diag ? pencil_dispatch2<W, 1>(ctx, Tr, a) : pencil_dispatch2<W, 0>(ctx, Tr, a);
}

int tri_pencil_solve(lsspg_ctx *ctx, const lsspg_tri *Tr, double *dx, const double *drhs, bool guarded)
{
    lsspg_tri *Tm = const_cast<lsspg_tri *>(Tr);
    if (Tm->p_seen_timeouts != ctx->tri_timeouts) {
        // a sweep was aborted by the watchdog: mailboxes may hold values; empty them
        LSSPG_TRY(vec_set(ctx, (int)std::max<long long>(Tr->p_mail_len, 1), Tr->p_mail, bits_to_double(kPenEmptyBits), false));
        LSSPG_CUDA(cudaMemsetAsync(Tr->d_counter, 0, sizeof(unsigned int) * 64, ctx->stream));
        Tm->p_seen_timeouts = ctx->tri_timeouts;
    }
    PencilArgs a;
    a.hdr = (const PencilHdr *)Tr->p_hdr; a.thr = (const PencilThread *)Tr->p_thr; a.ghost = (const PencilGhost *)Tr->p_ghost; a.outs = (const PencilOut *)Tr->p_outs;
    a.vals = Tr->p_vals; a.mail = Tr->p_mail; a.counter = Tr->d_counter;
    a.num_pencils = Tr->num_tiles; a.T = Tr->p_T; a.RS = Tr->p_RS; a.dir = Tr->p_dir;
    a.x = dx; a.rhs = drhs;
    a.stop = guarded ? ctx->d_flags + FLAG_STOP : nullptr;
    a.err = ctx->d_flags + FLAG_TRI_TIMEOUT;
    static int prof_on = -1;
    if (prof_on < 0) prof_on = getenv("LSSPG_TRI_PROF") ? 1 : 0;
    a.prof = nullptr;
    a.dbg = getenv("LSSPG_TRI_PENCIL_DBG") ? atoi(getenv("LSSPG_TRI_PENCIL_DBG")) : 0;
    if (prof_on) {
        if (!Tm->p_prof) LSSPG_CUDA(cudaMalloc(&Tm->p_prof, 64 * (size_t)std::max(Tr->num_tiles, 1)));
        a.prof = Tm->p_prof;
    }
#define PEN_CASE(w) \
    case w: return pencil_dispatch<w>(ctx, Tr, a);
    switch (Tr->p_W) {
        PEN_CASE(1) PEN_CASE(2) PEN_CASE(3) PEN_CASE(4) PEN_CASE(5) PEN_CASE(6) PEN_CASE(7) PEN_CASE(8)
    }
#undef PEN_CASE
    set_error("tri_pencil: unsupported row width %d", Tr->p_W);
    return 1;
}

int tri_pencil_upload(lsspg_ctx *ctx, const PencilHost &H, lsspg_tri *T)
{
    T->pencil = true;
    T->num_tiles = H.num_pencils;
    T->p_T = H.T; T->p_W = H.W; T->p_diag = (H.nv > H.W) ? 1 : 0; T->p_dir = H.dir;
    T->p_RS = pencil_ring_stride(H.T, H.max_ghost);
    T->p_holes = H.holes ? 1 : 0;
    T->p_mail_len = H.mail_len;
    T->p_seen_timeouts = ctx->tri_timeouts;
    T->p_pv = H.pv; T->p_pw = H.pw;
    for (int k = 0; k < 3; k++) { T->grid_dims[k] = H.grid_dims[k]; T->tile_dims[k] = 0; }
    T->tile_dims[0] = H.pv; T->tile_dims[1] = H.pw;
    T->max_tile_rows = H.T; T->num_tile_levels = H.max_steps;
    LSSPG_CUDA(cudaMalloc(&T->p_hdr, sizeof(PencilHdr) * std::max<size_t>(H.hdr.size(), 1)));
    LSSPG_CUDA(cudaMalloc(&T->p_thr, sizeof(PencilThread) * std::max<size_t>(H.thr.size(), 1)));
    LSSPG_CUDA(cudaMalloc(&T->p_ghost, sizeof(PencilGhost) * std::max<size_t>(H.ghost.size(), 1)));
    LSSPG_CUDA(cudaMalloc(&T->p_outs, sizeof(PencilOut) * std::max<size_t>(H.outs.size(), 1)));
    if (!H.outs.empty())
        LSSPG_CUDA(cudaMemcpyAsync(T->p_outs, H.outs.data(), sizeof(PencilOut) * H.outs.size(), cudaMemcpyHostToDevice, ctx->stream));
    T->p_own_last = H.own_last ? 1 : 0;
    LSSPG_CUDA(cudaMalloc(&T->p_vals, sizeof(double) * std::max<size_t>(H.vals.size(), 1)));
    LSSPG_CUDA(cudaMalloc(&T->p_mail, sizeof(double) * (size_t)std::max<long long>(H.mail_len, 1)));
    LSSPG_CUDA(cudaMemcpyAsync(T->p_hdr, H.hdr.data(), sizeof(PencilHdr) * H.hdr.size(), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(T->p_thr, H.thr.data(), sizeof(PencilThread) * H.thr.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!H.ghost.empty())
        LSSPG_CUDA(cudaMemcpyAsync(T->p_ghost, H.ghost.data(), sizeof(PencilGhost) * H.ghost.size(), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(T->p_vals, H.vals.data(), sizeof(double) * H.vals.size(), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_TRY(vec_set(ctx, (int)std::max<long long>(H.mail_len, 1), T->p_mail, bits_to_double(kPenEmptyBits), false));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

void tri_pencil_free(lsspg_tri *T)
{
    cudaFree(T->p_hdr);
    cudaFree(T->p_thr);
    cudaFree(T->p_ghost);
    cudaFree(T->p_outs);
    cudaFree(T->p_vals);
    cudaFree(T->p_mail);
    cudaFree(T->p_prof);
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

// CPU self-check of the pencil schedule for the test-suite (never called by a product path): builds the image
// lsspg_tri_analyse would upload and replays it on the host.  *applicable = 0 when the factor is not a lattice
// factor (the caller then gets the box or slice schedule).  info[8]: pencils, threads per pencil, most ghost lines,
// largest operand distance, slots per row, values per row, most steps, skew (s1 s2 t1 as digits).
int lsspg_debug_tri_walk_pencil_host(int which, int n, const int *hTp, const int *hTj, const double *hTx, double *hx,
                                     const double *hrhs, int *applicable, int *info)
{
    PencilHost H;
    const int rc = tri_pencil_build_host(which, n, hTp, hTj, hTx, 148, H);
    if (applicable) *applicable = (rc == 0);
    if (rc == 1) return 1;
    if (rc == 2) return 0;
    return tri_pencil_walk_host(H, hx, hrhs, info);
}


// LSSPG_TRI_PROF=1: per-pencil timers of the last sweep with this factor (8 words per pencil, ticket order: start ns,
// end ns, cycles total / waiting for ghost lanes / in the step barrier, steps, CTA, SM).  Returns the number of pencils.
int lsspg_debug_tri_pencil_prof(lsspg_ctx *ctx, const lsspg_tri *T, unsigned long long *out, int max_pencils)
{
    if (!T || !T->pencil || !T->p_prof) return 0;
    cudaStreamSynchronize(ctx->stream);
    const int np = std::min(T->num_tiles, max_pencils);
    if (cudaMemcpy(out, T->p_prof, 64 * (size_t)np, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return np;
}

}  // extern "C"
