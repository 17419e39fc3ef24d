"""Run under torchrun (one rank per GPU): the NCCL paths against a single-GPU run of the same problem on rank 0.
Direct sum: i-rows split over the ranks + in-place all-gather of positions each step. Barnes-Hut: replicated Morton
sort + tree, each rank walks / integrates its slice of the Morton order, all-gather of positions and velocities."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelnbody_b200 as P  # noqa: E402
from parallelnbody_b200 import ic, launch  # noqa: E402


def rel_l2(a, b):
    a = np.asarray(a, np.float64)[:, :3]; b = np.asarray(b, np.float64)[:, :3]
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_uid():      # one ncclUniqueId per communicator, created on rank 0 and broadcast out of band
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(P.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    ok = True
    # (method, bh_exchange, n, steps, position tolerance vs the single-GPU run)
    cases = ((P.METHOD_DIRECT, 0, 50_001, 10, 2e-6),       # same kernel, bit-identical in practice
             (P.METHOD_BARNES_HUT, 1, 200_003, 10, 1e-6),  # replicated tree: same tree, same groups
             (P.METHOD_BARNES_HUT, 0, 200_003, 10, 2e-5))  # domain split + LET: remote cells accepted more strictly
    for method, exch, n, steps, tol in cases:
        posm, vel = ic.plummer(n, seed=5)
        uid = fresh_uid()
        with P.OctreeSearch(method=method, eps=0.01, theta=0.3, device=local, rank=rank, world=world, nccl_unique_id=uid,
                            bh_exchange=exch) as s:
            s.SetBodies(posm, vel)
            ke, pe = s.Energy()
            s.CreateOctree()
            ids0 = s.LocalIds()
            acc0 = launch.combine_shares(s.Accelerations(), ids0, n, dist, device="cuda")
            s.Step(1e-3, steps)
            ids = s.LocalIds()
            full_p = launch.combine_shares(s.Positions(), ids, n, dist, device="cuda")
            full_v = launch.combine_shares(s.Velocities(), ids, n, dist, device="cuda")
            st = s.Stats()
            nloc = torch.tensor([float(st["n_local"])], device="cuda"); nmax = nloc.clone(); nmin = nloc.clone()
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX); dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            with P.OctreeSearch(method=method, eps=0.01, theta=0.3, device=local) as one:
                one.SetBodies(posm, vel)
                ke1, pe1 = one.Energy()
                one.CreateOctree()
                a1 = one.Accelerations()
                one.Step(1e-3, steps)
                p1, v1 = one.Positions(), one.Velocities()
            with P.OctreeSearch(method=P.METHOD_DIRECT, eps=0.01, device=local) as d:
                d.SetBodies(posm, vel); d.CreateOctree(); exact = d.Accelerations()
            ep, ev = rel_l2(full_p, p1), rel_l2(full_v, v1)
            err_multi, err_one = rel_l2(acc0, exact), rel_l2(a1, exact)
            good = (ep <= tol and ev <= 50 * tol and abs(ke - ke1) <= 1e-9 * abs(ke1) and abs(pe - pe1) <= 1e-6 * abs(pe1)
                    and err_multi <= err_one * 1.05 + 1e-6)
            print(f"method={method} exchange={exch} world={world} n={n}: pos rel-L2 {ep:.2e} vel rel-L2 {ev:.2e} | force error vs direct: "
                  f"{err_multi:.3e} (multi) {err_one:.3e} (1 GPU) | energy {ke + pe:.6e} vs {ke1 + pe1:.6e} | n_local {int(nmin.item())}..{int(nmax.item())} "
                  f"| ms/step force {st['ms_force'] / steps:.3f} build {st['ms_build'] / steps:.3f} comm {st['ms_comm'] / steps:.3f} -> {'ok' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0 and ok:
        print("MULTI-GPU CHECK OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
