"""bench.py --gpus N (N > 1): the row-sharded solve loop, one process per GPU (torchrun).

Weak scaling: every GPU owns a 256 x 256 x 256 slab of a 256 x 256 x (256 N) 7-point
Laplacian (BASELINE.json configs[1] per GPU; N = 8 is configs[3]'s 134 M rows with ILU(0)
instead of AMG).  CG with block-Jacobi ILU(0) (= the reference's blocked ILU, one block
per GPU), halo exchange of x by NCCL send/recv, dot products by NCCL all-reduce.

value = N x iterations / second: every iteration advances N shards of the per-GPU size,
so ideal weak scaling keeps iterations/s constant and `value` grows with N.
"""
import ctypes as C
import json
import os
import time

import numpy as np


def main(args, rank, world, local_rank):
    # stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...") on fd 1 from C,
    # so fd 1 is pointed at stderr for the run and the JSON line goes to the saved descriptor
    import sys
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as td

    from . import api, dist
    from . import generators as g
    from ._lib import check, lib

    torch.cuda.set_device(local_rank)
    td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = lib()
    N = args.grid
    powerlaw = args.workload == "idrs_powerlaw"
    strong = getattr(args, "scaling", "weak") == "strong"
    if strong:
        dims = (N, N, N)                      # ONE grid^3 problem, z-slabs of N / world planes
    elif getattr(args, "shape", "slab") == "cubic":
        f = [1, 1, 1]                         # the most cubic factorisation of the GPU count (8 -> 2 x 2 x 2)
        w = world
        while w > 1:
            p = next(q for q in (2, 3, 5, 7, w) if w % q == 0)
            f[f.index(min(f))] *= p
            w //= p
        dims = (N * f[0], N * f[1], N * f[2])
    else:
        dims = (N, N, N * world)
    n = args.pl_rows * world if powerlaw else dims[0] * dims[1] * dims[2]
    blk, r0, r1 = dist.block_rows(n, world, rank)
    bicg = args.workload in ("bicgstab_ilu0", "bicgstab_iluk1")   # iluk1: BASELINE.json configs[2] (C3), block-Jacobi ILU(1)
    op = getattr(args, "operator", None) or ("cd" if bicg else "lap")
    conv = (0.3, 0.2, 0.1) if op == "cd" else (0.0, 0.0, 0.0)
    solver = "idrs" if powerlaw else "bicgstab" if bicg else "cg"
    t0 = time.perf_counter()
    # BASELINE.json configs[4] (power-law CSR, IDRS(4)): --pl-rows rows per GPU, every rank generates its own block
    rows = g.powerlaw_rows(n, r0, r1) if powerlaw else g.stencil_7pt_rows(dims, r0, r1, conv=conv)
    shard = dist.make_shard(rows, n, world, rank)          # metadata exchange: all_gather_object
    t_gen = time.perf_counter() - t0
    ids = [dist.DeviceShard.unique_id() if rank == 0 else None]
    td.broadcast_object_list(ids, src=0)
    ctx = api.Context(local_rank)
    ctx.set_option(api.OPT_CHECK_EVERY, args.check_every)
    D = dist.DeviceShard(ctx, shard, ids[0])
    t0 = time.perf_counter()
    pcname = "block-Jacobi ILU(0)"
    if args.workload == "cg_non" or powerlaw:
        pc = api.Preconditioner.non(ctx, shard.n_owned)
        pcname = "no preconditioner"
    elif args.workload == "cg_amg":
        pc = api.Preconditioner.sxamg(ctx, shard.diag_block(), zero_guess=1, cf_order=args.amg_order)
        pcname = "block-Jacobi SX-AMG-style V-cycle (zero initial guess, cf_order %d)" % args.amg_order
    else:
        lvl = 1 if args.workload == "bicgstab_iluk1" else 0
        pcname = "block-Jacobi ILU(%d)" % lvl
        Lf, Uf = api.ilu_factor(shard.diag_block(), "iluk", level=lvl)
        pc = api.Preconditioner(ctx, "ilu", shard.n_owned, Lf, Uf)
    t_pc = time.perf_counter() - t0
    no, nc = shard.n_owned, shard.n_owned + shard.n_ghost
    # b = 1 as exam.cxx; the power-law operator has unit row sums (x = 1 would solve it at once): oscillating b
    hb_full = np.ones(nc)
    if powerlaw:
        hb_full[:no] = np.sin(np.arange(r0, r1) * 0.37) + 1.5
    b = ctx.upload(hb_full)
    x = ctx.zeros(nc)
    kw = dict(idrs=4) if powerlaw else {}

    def solve():
        check(L.lsspg_memset_zero(ctx.h, x.ptr, C.c_size_t(8 * nc)))
        return api.solve_device(ctx, solver, D.A, pc, b, x, maxit=3000, **kw)

    for _ in range(args.warmup):
        r = solve()
    from bench import ClockSampler
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.sync()
    td.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    ms = C.c_double()
    check(L.lsspg_timer_start(ctx.h, 0))
    its = 0
    for _ in range(args.steps):
        r = solve()
        its += r["nits"]
    check(L.lsspg_timer_stop(ctx.h, 0, C.byref(ms)))
    ctx.sync()
    td.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    total_ms = float(t[0])
    launches = ctx.launches - launches0
    # e2e: host b / x shards through the reference-facing call, copies inside the timed region
    hb, hx = np.ascontiguousarray(hb_full[:no]), np.zeros(no)
    e2e_its, e2e_s = 0, 0.0
    for step in range(1 + args.steps):
        hx[:] = 0.0
        ctx.sync()
        td.barrier()
        t0 = time.perf_counter()
        re = api.lssp_solver_solve(ctx, solver, D.A, pc, hb, hx, maxit=3000, **kw)
        _ = float(hx[no // 2])
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        td.all_reduce(dt, op=td.ReduceOp.MAX)
        if step >= 1:
            e2e_its += re["nits"]
            e2e_s += float(dt[0])
    clocks = sampler.finish() if sampler else None
    # roofline of the dominant streaming kernel: the sharded SpMV incl. its halo exchange (all ranks take part)
    y = ctx.empty(no)
    D.A.mv(api.MV_MXY, b, y)
    ctx.sync()
    td.barrier()
    tm = C.c_double()
    check(L.lsspg_timer_start(ctx.h, 1))
    for _ in range(20):
        D.A.mv(api.MV_MXY, b, y)
    check(L.lsspg_timer_stop(ctx.h, 1, C.byref(tm)))
    tmax = torch.tensor([tm.value / 20], dtype=torch.float64, device="cuda")
    td.all_reduce(tmax, op=td.ReduceOp.MAX)
    ms_spmv = float(tmax[0])
    # the preconditioner application (local to the rank: block-Jacobi), the dominant kernel of the iteration as at N = 1
    ms_pc = None
    if pcname != "no preconditioner":
        z = ctx.empty(nc)
        pc.apply(z, b)
        ctx.sync()
        td.barrier()
        check(L.lsspg_timer_start(ctx.h, 1))
        for _ in range(10):
            pc.apply(z, b)
        check(L.lsspg_timer_stop(ctx.h, 1, C.byref(tm)))
        tmax = torch.tensor([tm.value / 10], dtype=torch.float64, device="cuda")
        td.all_reduce(tmax, op=td.ReduceOp.MAX)
        ms_pc = float(tmax[0])
    if rank == 0:
        from bench import measured_peak
        peak, peak_kind = measured_peak()
        gbs = D.A.spmv_bytes / ms_spmv / 1e6
        roof_spmv = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                     "kernel": "SpMV on one rank's row block, halo exchange (pack kernel + NCCL send/recv) included",
                     "ms": ms_spmv, "bytes": D.A.spmv_bytes, "peak_kind": peak_kind, "per_gpu": True}
        ms_it = total_ms / its
        if ms_pc is not None and args.workload != "cg_amg":
            pgbs = pc.bytes / ms_pc / 1e6
            roof = {"bound": "hbm", "achieved": pgbs, "peak": peak, "unit": "GB/s", "frac": pgbs / peak, "traffic": None,
                    "kernel": "triangular sweep of the rank's block-Jacobi ILU (tri_pencil_kernel; 2 launches per application, "
                              "average of the two), the same kernel bench.py reports at N = 1",
                    "ms": ms_pc / 2, "bytes": pc.bytes / 2, "share_of_iteration": ms_pc * (2 if solver == "bicgstab" else 1) / ms_it,
                    "peak_kind": peak_kind, "per_gpu": True}
        else:
            roof = dict(roof_spmv)
        per_s = its / (total_ms / 1e3)
        value = per_s if strong else world * per_s
        line = {"metric": "%s_iterations_per_second" % args.workload, "value": value, "unit": "iter/s" if strong else "iter/s x n_gpus",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": ("power-law CSR n=%d (nnz %d on rank 0) IDRS(4), %s, row-sharded over %d GPUs "
                                        "(%d rows per GPU)" % (n, int(rows[0][-1]), pcname, world, args.pl_rows)) if powerlaw else
                                       "%s %dx%dx%d %s + %s, row-sharded (z-slabs) over %d GPUs "
                                       "(%d rows per GPU)" % ("lap3d" if op == "lap" else "cd3d", dims[0], dims[1], dims[2], solver.upper(),
                                                              pcname, world, no),
                           "n": n, "rows_per_gpu": no, "ghost_per_gpu": shard.n_ghost, "tol_rel": 1e-7,
                           "iterations_per_solve": r["nits"], "residual": r["residual"],
                           "iterations_per_second": its / (total_ms / 1e3),
                           "value_definition": "iterations/s of the one fixed-size problem" if strong else
                                               "n_gpus x iterations/s (each iteration advances n_gpus shards of the 1-GPU size)",
                           "allreduce": "one-shot peer-to-peer kernel (k_p2p_allreduce_fin)" if os.environ.get("LSSPG_P2P_ALLREDUCE", "1") != "0"
                                        else "ncclAllReduce + k_fin",
                           "preconditioner": "block-Jacobi (triangular sweeps / AMG hierarchy local to each GPU's row block, "
                                             "= reference blocked ILU with blk_size=ceil(n/P)); iteration counts depend on P "
                                             "(SURVEY.md App. A.5)",
                           "l2": "per-GPU working set (1.7 GB CSR + factors) far exceeds the 126 MB L2"},
                "ms_per_iteration": total_ms / its, "gpu_launches": int(launches), "clocks": clocks,
                "e2e": {"value": (1 if strong else world) * e2e_its / e2e_s, "unit": "iter/s" if strong else "iter/s x n_gpus",
                        "h2d_bytes_per_step": 16 * no * world, "d2h_bytes_per_step": 8 * no * world},
                "roofline": roof, "roofline_spmv": roof_spmv,
                "setup_s": {"generate_and_shard": t_gen, "pc_host_setup_and_upload": t_pc}}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    D.close()
    td.barrier()
    td.destroy_process_group()
    return 0
