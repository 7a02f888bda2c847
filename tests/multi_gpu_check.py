"""Multi-GPU global fit check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py [n_total]

Every rank holds a contiguous shard of ONE sample set.  The fit runs twice:
  * host loop + NCCL all-reduce of the 10 sums per evaluation (DRIVE_HOST)
  * the persistent kernel with the fused in-kernel peer-memory exchange (DRIVE_PERSISTENT)
and both must (a) give bit-identical p / info on every rank, (b) agree with the single-GPU fit of the
whole set on rank 0 and with the CPU oracle within the parity tolerances (1e-4 params, 1e-6 cost).
Prints "MULTI_GPU_OK ..." on rank 0; any failure raises (non-zero exit).  Used by
tests/test_gpu_multi.py and by hand under gpurun --gpus N.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist

    import oracle_lib as O
    import synth
    from brdf_b200 import api as A

    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 200_001
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = A.Context(local)

    ids = [A.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.peer_export())
    ctx.peer_attach(handles, rank, world)

    c, td, th, x = synth.samples(n_total, seed=77)
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world   # ragged shards on purpose
    s = ctx.upload(c[lo:hi], td[lo:hi], x[lo:hi], A.BLINN_PHONG)

    results = {}
    for name, drive in (("nccl-host", A.DRIVE_HOST), ("peer-persistent", A.DRIVE_PERSISTENT)):
        for preset_name, preset in (("REF_GLOBAL", A.REF_GLOBAL), ("REF_PERFACE", A.REF_PERFACE)):
            dist.barrier()
            t0 = time.perf_counter()
            ret, p, info = ctx.fit_global(s, preset, drive=drive)
            dt = time.perf_counter() - t0
            blob = np.concatenate([[ret], p, info])
            all_blobs = [None] * world
            dist.all_gather_object(all_blobs, blob.tobytes())
            assert all(b == all_blobs[0] for b in all_blobs), "%s/%s: ranks disagree" % (name, preset_name)
            results[(name, preset_name)] = (ret, p, info, dt)

    # ---- re-attach: a second (third) session over the SAME exchange buffers.  The buffers still hold the cells of the
    # earlier sessions; tags never go back (the export record carries the tag reached), so none of them can be taken
    # for a fresh one.  Each session runs far more exchanges than the two parities a buffer holds.
    want = np.concatenate([[results[("peer-persistent", "REF_GLOBAL")][0]], results[("peer-persistent", "REF_GLOBAL")][1],
                           results[("peer-persistent", "REF_GLOBAL")][2]]).tobytes()
    for session in range(2):
        ctx.peer_detach()
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_export())
        ctx.peer_attach(handles, rank, world)      # no barrier, no memset: the first fit may start at once
        for _ in range(2):
            ret, p, info = ctx.fit_global(s, A.REF_GLOBAL, drive=A.DRIVE_PERSISTENT)
            assert np.concatenate([[ret], p, info]).tobytes() == want, "re-attached session %d differs" % session
    # ---- covariance from a context that only has peer buffers (no NCCL communicator): the global sample count
    # travels through the fit kernel's own exchange
    lone = A.Context(local)
    handles = [None] * world
    dist.all_gather_object(handles, lone.peer_export())
    lone.peer_attach(handles, rank, world)
    s_lone = lone.upload(c[lo:hi], td[lo:hi], x[lo:hi], A.BLINN_PHONG)
    ret_l, p_l, info_l, covar_l = lone.fit_global(s_lone, A.REF_GLOBAL, want_covar=True)
    assert np.concatenate([[ret_l], p_l, info_l]).tobytes() == want
    s_lone.free()
    dist.barrier()
    lone.close()

    if rank == 0:
        single = A.Context(local)
        s1 = single.upload(c, td, x, A.BLINN_PHONG)
        _, _, _, covar_1 = single.fit_global(s1, A.REF_GLOBAL, want_covar=True)
        np.testing.assert_allclose(covar_l, covar_1, rtol=1e-6)
        print("re-attached twice, covariance without a communicator: ok", flush=True)
        for preset_name, preset, opreset in (("REF_GLOBAL", A.REF_GLOBAL, O.REF_GLOBAL), ("REF_PERFACE", A.REF_PERFACE, O.REF_PERFACE)):
            r1, p1, i1 = single.fit_global(s1, preset)
            wret, wp, winfo = O.brdf_fit(O.oracle(), "oracle_", c, td, th, x, 1, opreset)
            for name in ("nccl-host", "peer-persistent"):
                ret, p, info, dt = results[(name, preset_name)]
                assert (ret >= 0) == (wret >= 0) == (r1 >= 0)
                np.testing.assert_allclose(p, p1, rtol=1e-4, err_msg=name)
                np.testing.assert_allclose(p, wp, rtol=1e-4, err_msg=name)
                np.testing.assert_allclose(info[1], winfo[1], rtol=1e-6, err_msg=name)
                if int(info[6]) not in (3, 5) and int(winfo[6]) not in (3, 5):
                    assert int(info[6]) == int(winfo[6]), (name, info[6], winfo[6])
                print("%-16s %-11s ranks=%d n=%d iters=%d nfev=%d p=%s cost=%.12g  %.2f ms" %
                      (name, preset_name, world, n_total, info[5], info[7], p, info[1], dt * 1e3), flush=True)
        s1.free()
        single.close()

    # ---- the two modes that shard with NO data-path collective ----
    # gather: views are independent units -> rank r takes views r, r+world, ...; the concatenation in view order
    # must be the single-GPU gather of all views, byte for byte (results travel through the launcher only)
    import scene_lib as S
    W, H = 320, 240
    V, F = S.height_field(40, 30, seed=5)
    imgs, dark = S.random_images(16, W, H, seed=6)
    cams = np.array([S.look_at_camera((40.0 * np.cos(a), 30.0 * np.sin(a), 250.0), (0.0, 0.0, 0.0), f=400.0, cx=160.0, cy=120.0)
                     for a in np.linspace(0.0, 3.0, 2 * world + 1)])
    sc = ctx.scene(V, F, imgs, dark)
    mine = list(range(rank, len(cams), world))
    g = sc.gather(cams[mine])
    off = np.concatenate([[0], np.cumsum(g["nfit_cam"])])
    parts = {v: {k: np.ascontiguousarray(g[k][off[j]:off[j + 1]]).tobytes() for k in ("fit_face", "fit_pixel", "phi", "thetaDash", "theta")}
             for j, v in enumerate(mine)}
    for j, v in enumerate(mine):
        parts[v]["map"] = g["maps"][j].tobytes()
        parts[v]["I"] = np.ascontiguousarray(g["I"][:, off[j]:off[j + 1]]).tobytes()
    all_parts = [None] * world
    dist.all_gather_object(all_parts, parts)
    # batched fits: fit ids [lo, hi) per rank; the union must be the single-GPU batch
    nfit = 1000 * world + 7
    blo, bhi = rank * nfit // world, (rank + 1) * nfit // world
    b = ctx.batch_synth(bhi - blo, 16, seed=9, first_fit=blo)
    b.fit(A.REF_PERFACE)
    bp, binfo, bret = b.results()
    all_b = [None] * world
    dist.all_gather_object(all_b, (bp.tobytes(), binfo.tobytes(), bret.tobytes()))
    if rank == 0:
        merged = {}
        for d in all_parts:
            merged.update(d)
        gg = sc.gather(cams)
        o = np.concatenate([[0], np.cumsum(gg["nfit_cam"])])
        for v in range(len(cams)):
            for k in ("fit_face", "fit_pixel", "phi", "thetaDash", "theta"):
                assert merged[v][k] == np.ascontiguousarray(gg[k][o[v]:o[v + 1]]).tobytes(), (v, k)
            assert merged[v]["map"] == gg["maps"][v].tobytes() and merged[v]["I"] == np.ascontiguousarray(gg["I"][:, o[v]:o[v + 1]]).tobytes()
        print("gather sharded by view: %d views over %d ranks, %d fits, byte-identical to one GPU" % (len(cams), world, int(o[-1])), flush=True)
        b1 = ctx.batch_synth(nfit, 16, seed=9)
        b1.fit(A.REF_PERFACE)
        p1, i1, r1 = b1.results()
        assert b"".join(x[0] for x in all_b) == p1.tobytes() and b"".join(x[1] for x in all_b) == i1.tobytes()
        assert b"".join(x[2] for x in all_b) == r1.tobytes()
        print("batched fits sharded by fit id: %d fits over %d ranks, byte-identical to one GPU" % (nfit, world), flush=True)
        b1.free()
        print("MULTI_GPU_OK world=%d" % world, flush=True)
    b.free()
    sc.free()
    dist.barrier()
    s.free()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
