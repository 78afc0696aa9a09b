"""N>1 path on the CPU: world_size-2 gloo run of the seed-batch sharding + the one gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_generate(unit: int) -> torch.Tensor:
    from diffusionspatialcontrol_b200.distributed import unit_noise

    x = unit_noise(unit, 2, (4, 8, 8))
    return torch.tanh(x) * (unit + 1)  # deterministic function of the unit alone


def _worker(rank, world, port, n_units, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusionspatialcontrol_b200.distributed import run_sharded

    out = run_sharded(_fake_generate, n_units)
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def _worker_too_few(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusionspatialcontrol_b200.distributed import run_sharded

    calls = []
    try:
        run_sharded(lambda u: calls.append(u) or _fake_generate(u), 1)
        verdict = "no error"
    except ValueError as e:
        verdict = f"ValueError after {len(calls)} generate calls: {e}"
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write(verdict)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_units", [2, 5])
def test_two_ranks_gather_is_identical_to_one_rank(tmp_path, n_units):
    from diffusionspatialcontrol_b200.distributed import run_sharded, seeds_of_unit, units_for_rank

    assert units_for_rank(8, 1, 4) == [1, 5] and units_for_rank(3, 2, 4) == [2] and units_for_rank(2, 3, 4) == []
    assert seeds_of_unit(3, 8) == list(range(24, 32))
    single = run_sharded(_fake_generate, n_units)
    mp.spawn(_worker, args=(2, _free_port(), n_units, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a, b) and torch.equal(a, single)  # bit-identical at 1 and 2 ranks


def test_noise_depends_on_seed_only():
    from diffusionspatialcontrol_b200.distributed import unit_noise

    a = unit_noise(1, 8, (4, 8, 8))
    b = unit_noise(0, 16, (4, 8, 8))
    assert torch.equal(a, b[8:])


def test_fewer_units_than_ranks_raises_on_every_rank_before_any_work(tmp_path):
    """a rank without units used to raise only after its peers had entered the all_gather (they hung).  The
    check now runs first, with the same verdict on every rank, before any generate() call or collective."""
    mp.spawn(_worker_too_few, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        txt = (tmp_path / f"r{r}.txt").read_text()
        assert txt.startswith("ValueError after 0 generate calls"), txt
