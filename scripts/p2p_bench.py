#!/usr/bin/env python
"""Latency of the multi-GPU primitives (torchrun, one rank per GPU): interface exchange on
129^3 / 17^3 / 3^3 boxes, fused dot+all-reduce, plain all-reduce — peer windows vs NCCL
(UG4B200_P2P=0).  Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from ugcore_b200 import capi, dist as ugdist, solver as S
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    S.host_init(lr, None)
    ugdist.nccl_bootstrap(dist)
    p2p = ugdist.p2p_bootstrap(dist)
    dev, ctx = capi.dev, S.host_ctx()
    part = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    out = {"world": world, "p2p": bool(p2p)}
    e0, e1 = C.c_void_p(), C.c_void_p()
    dev.ug4b200_event_create(ctx, C.byref(e0)); dev.ug4b200_event_create(ctx, C.byref(e1))

    def timeit(fn, reps=200):
        for _ in range(20):
            fn()
        dev.ug4b200_sync(ctx)
        dist.barrier(); torch.cuda.synchronize()
        dev.ug4b200_event_record(ctx, e0)
        for _ in range(reps):
            fn()
        dev.ug4b200_event_record(ctx, e1)
        dev.ug4b200_event_sync(ctx, e1)
        ms = C.c_float()
        dev.ug4b200_event_elapsed_ms(ctx, e0, e1, C.byref(ms))
        return ms.value / reps * 1e3

    refs = 7
    prob = ugdist.local_problem(refs, part, rank)
    keep = []
    for lev in (7, 4, 1):
        ranks, ptr, idx = ugdist.interfaces(prob, lev)
        n = prob.matrix(lev).nrows
        I = C.c_void_p()
        capi.check(dev.ug4b200_interface_create(ctx, ranks.size, ranks.ctypes.data_as(C.c_void_p), ptr.ctypes.data_as(C.c_void_p),
                                                idx.ctypes.data_as(C.c_void_p), n, C.byref(I)), ctx)
        capi.check(dev.ug4b200_interface_commit(ctx, I), ctx)
        v = S.DeviceBuffer.from_numpy(np.zeros(n))
        keep.append((I, v))
        out[f"exchange_lev{lev}_us"] = timeit(lambda: dev.ug4b200_additive_to_consistent(ctx, I, v.ptr, 1))
        out[f"exchange_lev{lev}_entries"] = int(idx.size)
    n = prob.matrix(7).nrows
    a = S.DeviceBuffer.from_numpy(np.ones(n)); res = S.DeviceBuffer.from_numpy(np.zeros(8))
    fin = capi.Fin(capi.FIN_STORE, res.ptr, None, None, None)
    scratch = C.c_void_p(res.ptr.value + 8)
    out["dot_allreduce_us"] = timeit(lambda: dev.ug4b200_vec_dot_allreduce_ds(ctx, n, a.ptr, a.ptr, fin, scratch))
    out["dot_local_us"] = timeit(lambda: dev.ug4b200_vec_dot_ds(ctx, n, a.ptr, a.ptr, fin))
    out["allreduce_8_us"] = timeit(lambda: dev.ug4b200_allreduce_sum(ctx, res.ptr, 8))
    small = S.DeviceBuffer.from_numpy(np.ones(64))
    out["dot_allreduce_small_us"] = timeit(lambda: dev.ug4b200_vec_dot_allreduce_ds(ctx, 64, small.ptr, small.ptr, fin, scratch))
    out["vec_set_small_us"] = timeit(lambda: dev.ug4b200_vec_set(ctx, 64, small.ptr, 0.0))
    if rank == 0:
        print("P2PBENCH " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
