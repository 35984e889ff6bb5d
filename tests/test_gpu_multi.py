"""Real multi-GPU run of the sharded evaluator (NCCL): needs >= 2 GPUs, skipped otherwise.
The 1-GPU suite already proves shard-count independence by emulating the ranks on one device
(test_gpu_parity.py::test_slab_decomposition_is_bit_identical, ::test_tile_shard_decomposition_is_bit_identical);
this test closes the loop through
torch.distributed: every rank must return the single-GPU answer bit for bit."""
import json
import os
import socket
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from irrotational_warp_lab_b200 import distributed, sweeps
        out = {}
        for n, dtype in ((96, "float64"), (200, "float64"), (100, "float32")):
            for shard in ("tiles", "slab"):       # interleaved tile shards (default) and slabs of axis 0
                run = distributed.compute_energy_3d_sharded(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n, dtype=dtype, shard=shard)
                assert run["engine"]["shard"] == shard
                extra = run["_e_at_r"](23.5)                      # collective relaunch with an unseen radius
                out[f"{n}_{dtype}_{shard}"] = [run["energies"]["e_pos"].hex(), run["energies"]["e_neg"].hex(), run["counts"]["neg"],
                                               run["_e_at_r"](40.0)[0].hex(), extra[0].hex(), extra[1].hex()]
        sw = sweeps.sweep_velocity(mode="3d", rho=5.0, sigma=4.0, v_values=[1.0, 1.5, 2.0, 2.5, 3.0], extent=66.0,
                                   tail_r1_mult=8.0, tail_r2_mult=12.0, n=64, distributed=True)   # collective: points dealt to ranks
        out["sweep"] = [r["e_pos_inf"].hex() for r in sw["sweep_results"]]
        json.dump(out, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ranks = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(world)]
    assert ranks[0] == ranks[1]
    from irrotational_warp_lab_b200 import backend, sweeps
    for n, dtype in ((96, "float64"), (200, "float64"), (100, "float32")):
        one = backend.compute_energy_3d(rho=5.0, sigma=4.0, v=1.0, extent=66.0, n=n, dtype=dtype)
        extra = one["_e_at_r"](23.5)
        for shard in ("tiles", "slab"):
            assert ranks[0][f"{n}_{dtype}_{shard}"] == [one["energies"]["e_pos"].hex(), one["energies"]["e_neg"].hex(), one["counts"]["neg"],
                                                        one["_e_at_r"](40.0)[0].hex(), extra[0].hex(), extra[1].hex()]
    sw = sweeps.sweep_velocity(mode="3d", rho=5.0, sigma=4.0, v_values=[1.0, 1.5, 2.0, 2.5, 3.0], extent=66.0,
                               tail_r1_mult=8.0, tail_r2_mult=12.0, n=64)
    assert ranks[0]["sweep"] == [r["e_pos_inf"].hex() for r in sw["sweep_results"]]
