// scalars.cuh -- device-resident Krylov scalars.
//
// The drivers keep alpha/beta/omega/rho on the device (ctx->d_scal) so that one
// iteration needs at most one device->host read (the residual).  Kernels take
// their coefficients as `Coef` (immediate or slot in the slab) and reducing
// kernels run a tiny `FinProg` in their last CTA to derive the next scalars
// (e.g. alpha = rho / (p.q)) with the same IEEE double operations, in the same
// order, as the reference's host code.
#pragma once

namespace lsspg {

struct Coef {
    int slot;     // >= 0: value = scal[slot]; < 0: value = imm
    int neg;      // negate after load (exact)
    double imm;
};

inline Coef coef_imm(double v) { Coef c; c.slot = -1; c.neg = 0; c.imm = v; return c; }
inline Coef coef_slot(int s, bool neg = false) { Coef c; c.slot = s; c.neg = neg ? 1 : 0; c.imm = 0.0; return c; }

enum FinOpCode : int {
    FIN_DIV = 0,     // s[d] = s[a] / s[b]
    FIN_MUL = 1,     // s[d] = s[a] * s[b]
    FIN_SUB = 2,     // s[d] = s[a] - s[b]
    FIN_ADD = 3,     // s[d] = s[a] + s[b]
    FIN_SQRT = 4,    // s[d] = sqrt(s[a])
    FIN_NEG = 5,     // s[d] = -s[a]
    FIN_COPY = 6,    // s[d] = s[a]
    FIN_FLAG_LE = 7, // if (s[a] <= s[b]) flags[d] = v
    FIN_FLAG_EQ0 = 8 // if (s[a] == 0)   flags[d] = v
};

struct FinOp {
    int op, d, a, b, v;
};

constexpr int kMaxFinOps = 12;

struct FinProg {
    int n = 0;
    FinOp ops[kMaxFinOps];
    FinProg &add(int op, int d, int a, int b = 0, int v = 1)
    {
        if (n < kMaxFinOps) {
            ops[n].op = op; ops[n].d = d; ops[n].a = a; ops[n].b = b; ops[n].v = v;
            n++;
        }
        return *this;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ double coef_get(const Coef &c, const double *scal)
{
    double v = (c.slot >= 0) ? scal[c.slot] : c.imm;
    return c.neg ? -v : v;
}

__device__ __forceinline__ void fin_run(const FinProg &p, double *s, int *flags)
{
    for (int i = 0; i < p.n; i++) {
        const FinOp o = p.ops[i];
        switch (o.op) {
            case FIN_DIV: s[o.d] = s[o.a] / s[o.b]; break;
            case FIN_MUL: s[o.d] = s[o.a] * s[o.b]; break;
            case FIN_SUB: s[o.d] = s[o.a] - s[o.b]; break;
            case FIN_ADD: s[o.d] = s[o.a] + s[o.b]; break;
            case FIN_SQRT: s[o.d] = sqrt(s[o.a]); break;
            case FIN_NEG: s[o.d] = -s[o.a]; break;
            case FIN_COPY: s[o.d] = s[o.a]; break;
            case FIN_FLAG_LE: if (s[o.a] <= s[o.b]) flags[o.d] = o.v; break;
            case FIN_FLAG_EQ0: if (s[o.a] == 0.0) flags[o.d] = o.v; break;
        }
    }
}
#endif

}  // namespace lsspg
