"""Multi-GPU correctness + timing check of eigen_b200.parallelize.DistGemm (run under torchrun, one rank per GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import eigen_b200  # noqa: E402
from eigen_b200 import parallelize  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
shapes = [("d", 3000, 2500, 4100, 0.7, 1.3), ("d", 8192, 8192, 8192, 1.0, 1.0), ("s", 4096, 4096, 4096, 1.0, 1.0),
          ("z", 2048, 2048, 2048, 1.0, 1.0)]
for (t, m, n, k, alpha, beta) in shapes:
    dt = {"s": torch.float32, "d": torch.float64, "z": torch.complex128}[t]
    job = parallelize.DistGemm(t, m, n, k, alpha, beta)
    if rank == 0:
        g = torch.Generator(device="cuda").manual_seed(1)
        if t == "z":
            A = torch.view_as_complex(torch.rand(k, m, 2, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
            B = torch.view_as_complex(torch.rand(n, k, 2, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
        else:
            A = torch.rand(k, m, dtype=dt, device="cuda", generator=g) * 2 - 1
            B = torch.rand(n, k, dtype=dt, device="cuda", generator=g) * 2 - 1
        C0 = torch.ones(n, m, dtype=dt, device="cuda")
        want = C0.clone()
        eigen_b200.gemm_dev(t, "N", "N", m, n, k, alpha, A, m, B, k, beta, want, m)
        torch.cuda.synchronize()
    else:
        A = B = C0 = None
    for it in range(3):
        C = C0.clone() if rank == 0 else None
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        job.run(A, B, C)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        err = ((C - want).abs().max() / want.abs().max()).item()
        fl = (8.0 if t in "cz" else 2.0) * m * n * k
        print("dist %sgemm %dx%dx%d world=%d grid=%s: %.3f ms %.2f TF, max rel diff vs 1-GPU %.3e" % (
            t, m, n, k, world, (job.pr, job.pc), ms.item(), fl / ms.item() / 1e9, err), flush=True)
        assert err < (1e-4 if t == "s" else 1e-12), err
    del job
dist.barrier()
dist.destroy_process_group()
