"""world_size 2 and 3 on CPU (gloo): the element partition, the MPI-interface containers and the halo
exchange protocol, driven through the oracle's distributed RHS (dg_2d_parallel.jl:453-563).  The
reference asserts that MPI runs reproduce the serial results (test/test_mpi_p4est_3d.jl:9-12); here the
gathered distributed du must equal the serial du bit for bit (the same flux is evaluated with the same
operands on both sides of a shared face)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import trixi_b200 as T
    from elixirs import ELIXIRS, EXTRA
    from trixi_b200.parallel import HostHaloExchange, allreduce_min
    ex = ELIXIRS.get(name) or EXTRA[name]
    semi = ex.semi()
    psemi = ex.build()
    psemi.__init__(semi.mesh, semi.equations, semi.initial_condition, semi.solver, source_terms=semi.source_terms,
                   boundary_conditions=semi.boundary_conditions, rank=rank, world_size=world, comm=dist)
    ob = oracle.OracleBackend(psemi, num_threads=1)
    ob.set_halo_exchange(HostHaloExchange(psemi, dist).exchange)
    u = T.compute_coefficients(0.0, psemi)
    du = np.empty_like(u)
    ob.rhs_host(du, u, 0.3)
    # one full RK step and the distributed CFL step size
    ob.upload(0, u)
    dt_local = ob.max_dt()
    dt = allreduce_min(dt_local, dist)
    alg = T.CarpenterKennedy2N54()
    ob.step_2n(0.0, 0.5 * dt, alg.a, alg.b, alg.c)
    l2, linf = T.calc_error_norms(ob.download(0), 0.5 * dt, psemi)  # reductions over ranks
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), du=du, u1=ob.download(0), dt=dt, l2=l2, linf=linf,
             first=psemi.cache.first_element, last=psemi.cache.last_element)
    dist.barrier()
    dist.destroy_process_group()


_CASES = [(name, world) for name in ("tree_3d_euler_ec", "tree_2d_euler_source_terms_nonperiodic", "tree_3d_mhd_ec",
                                      "p4est_3d_euler_source_terms_nonperiodic", "p4est_3d_curved_level1")
          for world in (2, 3)]
# mortars that straddle ranks (MPI mortars): with 3 ranks some ranks own small elements of a mortar only
_CASES += [("tree_2d_advection_mortar", 2), ("tree_3d_euler_mortar", 3), ("tree_3d_mhd_alfven_wave_mortar", 3),
           ("p4est_2d_advection_nonconforming_flag", 2), ("p4est_2d_advection_nonconforming_flag", 3),
           ("p4est_3d_nonconforming_curved_ec", 3), ("p4est_3d_nonconforming_curved_weak_form_nonperiodic", 2),
           ("p4est_3d_mhd_alfven_wave_nonconforming", 3)]


@pytest.mark.parametrize("name,world", _CASES)
def test_distributed_oracle_equals_serial(world, name, tmp_path, oracle_module):
    import trixi_b200 as T
    from elixirs import ELIXIRS, EXTRA
    ELIXIRS = {**ELIXIRS, **EXTRA}
    port = 29500 + (os.getpid() + world * 7 + len(name)) % 2000
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    semi = ELIXIRS[name].semi()
    ob = oracle_module.OracleBackend(semi, num_threads=2)
    u = T.compute_coefficients(0.0, semi)
    du = np.empty_like(u)
    ob.rhs_host(du, u, 0.3)
    ob.upload(0, u)
    dt = ob.max_dt()
    alg = T.CarpenterKennedy2N54()
    ob.step_2n(0.0, 0.5 * dt, alg.a, alg.b, alg.c)
    u1 = ob.download(0).reshape(u.shape, order="F")
    l2, linf = T.calc_error_norms(u1, 0.5 * dt, semi)
    covered = 0
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        first, last = int(z["first"]), int(z["last"])
        covered += last - first
        assert float(z["dt"]) == dt
        np.testing.assert_allclose(z["l2"], l2, rtol=1e-12, atol=1e-15)  # psi of MHD is round-off only
        np.testing.assert_allclose(z["linf"], linf, rtol=1e-12, atol=1e-15)
        if name.startswith("p4est"):
            # each rank evaluates a shared face with the normal of its own element
            # (dgsem_p4est/dg_3d_parallel.jl:262-266): equal to rounding of the metric terms, not bit for bit
            # On the refined forests the absolute rounding of the metric terms (1e-15 on normals of size h/2) is
            # amplified by inverse_jacobian * inverse_weights (2300 * 10 at level 4, polydeg 4 in 2D; more in 3D); the MPI mortars
            # themselves use the small elements' normals on every rank and agree bit for bit.
            tol = 2e-10 if "nonconforming" in name else 1e-13
            np.testing.assert_allclose(z["du"], du[..., first:last], rtol=0, atol=tol * np.abs(du).max())
        else:
            np.testing.assert_array_equal(z["du"], du[..., first:last])
        # the stage update of the distributed driver is NumPy (no FMA contraction): 1-ulp differences
        np.testing.assert_allclose(z["u1"].reshape(du[..., first:last].shape, order="F"), u1[..., first:last],
                                   rtol=0, atol=1e-12 if "nonconforming" in name else 1e-14)
    assert covered == semi.nelements


def test_partition_is_contiguous_and_balanced():
    from trixi_b200.containers import owner_of, partition_cells
    for n, w in [(512, 8), (100, 3), (7, 7), (1000, 6)]:
        ranges = [partition_cells(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
        owners = owner_of(np.arange(n), n, w)
        for r, (a, b) in enumerate(ranges):
            assert np.all(owners[a:b] == r)
    with pytest.raises(ValueError):
        partition_cells(3, 0, 4)


def test_bench_weak_scaling_box_builds_on_every_rank():
    """bench.py's N > 1 workload: Morton-ordered CartesianBoxMesh blocks; every rank's containers and
    descriptor assemble (no GPU needed) and the shared faces pair up."""
    import bench
    world = 4
    semis = [bench.make_semi(2, rank=r, world=world) for r in range(world)]
    assert sum(s.nelements for s in semis) == semis[0].mesh.ncells
    for s in semis:
        s.descriptor()
        assert s.cache.mortars.nmortars == 0 and s.cache.boundaries.nboundaries == 0
    # every MPI interface appears on exactly two ranks with opposite sides
    gids = np.concatenate([s.cache.mpi_interfaces.global_interface_ids for s in semis])
    sides = np.concatenate([s.cache.mpi_interfaces.local_sides for s in semis])
    order = np.argsort(gids, kind="stable")
    g, sd = gids[order], sides[order]
    assert g.size % 2 == 0 and np.array_equal(g[0::2], g[1::2]) and np.all(sd[0::2] + sd[1::2] == 3)
