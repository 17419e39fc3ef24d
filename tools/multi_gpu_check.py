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
from parallelnbody_b200 import ic  # noqa: E402


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
    for method, n, steps, tol in ((P.METHOD_DIRECT, 50_001, 10, 2e-6), (P.METHOD_BARNES_HUT, 200_003, 10, 1e-6)):
        posm, vel = ic.plummer(n, seed=5)
        uid = fresh_uid()
        with P.OctreeSearch(method=method, eps=0.01, theta=0.3, device=local, rank=rank, world=world, nccl_unique_id=uid) as s:
            s.SetBodies(posm, vel)
            ke, pe = s.Energy()
            s.Step(1e-3, steps)
            mine_p, mine_v = s.Positions(), s.Velocities()
            ids = s.LocalIds()
            st = s.Stats()
        # combine the ranks' disjoint shares
        full_p = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        full_v = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
        idt = torch.from_numpy(ids).cuda()
        full_p[idt] = torch.from_numpy(mine_p[ids]).cuda()
        full_v[idt] = torch.from_numpy(mine_v[ids]).cuda()
        cnt[idt] += 1
        for x in (full_p, full_v, cnt):
            dist.all_reduce(x)
        if rank == 0:
            with P.OctreeSearch(method=method, eps=0.01, theta=0.3, device=local) as one:
                one.SetBodies(posm, vel)
                ke1, pe1 = one.Energy()
                one.Step(1e-3, steps)
                p1, v1 = one.Positions(), one.Velocities()
            ep, ev = rel_l2(full_p.cpu().numpy(), p1), rel_l2(full_v.cpu().numpy(), v1)
            good = bool((cnt == 1).all().item()) and ep <= tol and ev <= 50 * tol and abs(ke - ke1) <= 1e-9 * abs(ke1) and abs(pe - pe1) <= 1e-6 * abs(pe1)
            print(f"method={method} world={world} n={n}: pos rel-L2 {ep:.2e} vel rel-L2 {ev:.2e} energy {ke + pe:.6e} vs {ke1 + pe1:.6e} "
                  f"ms_comm {st['ms_comm']:.3f} ms_force {st['ms_force']:.3f} -> {'ok' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0 and ok:
        print("MULTI-GPU CHECK OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
