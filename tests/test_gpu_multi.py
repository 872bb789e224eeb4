"""Two processes, two GPUs: the halo exchange across processes (CUDA IPC mapped receive buffers).  Needs
`gpurun --gpus 2`; skipped on a single-GPU box (the in-process multi-rank test in test_gpu_parity.py covers
the same kernels there)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, name="tree_3d_euler_ec"):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import trixi_b200 as T
    from elixirs import ELIXIRS
    ex = ELIXIRS[name]
    base = ex.semi()
    semi = T.SemidiscretizationHyperbolic(base.mesh, base.equations, base.initial_condition, base.solver,
                                          source_terms=base.source_terms,
                                          boundary_conditions=base.boundary_conditions,
                                          rank=rank, world_size=world, comm=dist, device=rank)
    sol, l2, linf = ex.run(semi)  # full elixir run: StepsizeCallback allreduce + AnalysisCallback reductions
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), l2=l2, linf=linf, steps=sol.integrator.iter)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tree_3d_euler_ec", "p4est_3d_euler_source_terms_nonperiodic", "tree_3d_mhd_ec",
                                  "tree_2d_euler_vortex_shockcapturing",  # test/test_mpi_tree.jl:337-356
                                  # MPI mortars: test/test_mpi_tree.jl (elixir_advection_mortar.jl) and
                                  # test/test_mpi_p4est_2d.jl:33-45 assert the serial values for the MPI runs
                                  "tree_2d_advection_mortar", "p4est_2d_advection_nonconforming_flag"])
def test_two_gpu_run_reproduces_golden(name, tmp_path):
    """like test/test_mpi_p4est_3d.jl: the distributed run reproduces the serial golden values"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from elixirs import ELIXIRS
    mp.spawn(_worker, args=(2, 29733 + len(name), str(tmp_path), name), nprocs=2, join=True)
    ex = ELIXIRS[name]
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        ex.check(z["l2"], z["linf"])
