"""CPU (gloo, world_size 2) tests of the multi-GPU host logic: shard arithmetic and the tile
all-gather used by the large-swarm mode."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from abm_b200 import multigpu as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replicate_shards_partition_the_batch():
    for B, world in [(1024, 8), (10, 4), (3, 8), (7, 2)]:
        spans = [mg.replicate_shard(B, world, r) for r in range(world)]
        assert sum(c for _, c in spans) == B
        pos = 0
        for b, c in spans:
            assert b == pos
            pos += c


def test_agent_tiles():
    assert [mg.agent_tile(65536, 8, r) for r in (0, 7)] == [(0, 8192), (57344, 8192)]
    try:
        mg.agent_tile(10, 4, 0)
        assert False
    except ValueError:
        pass


def test_cyclic_slots():
    s = [mg.cyclic_slots(2048, 4, r) for r in range(4)]
    assert sorted(np.concatenate(s).tolist()) == list(range(2048))
    assert s[1][:3].tolist() == [128, 129, 130] and s[1][128] == 128 * 5
    try:
        mg.cyclic_slots(1000, 4, 0)
        assert False
    except ValueError:
        pass


def test_gloo_tile_allgather_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from abm_b200 import multigpu as mg
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        N = 64
        begin, count = mg.agent_tile(N, world, rank)
        full = torch.arange(N * 4, dtype=torch.float32).view(N, 4)
        table = torch.zeros(N, 4)
        table[begin:begin + count] = full[begin:begin + count] * (1.0)      # this rank's tile of records
        out = mg.gather_tiles(table[begin:begin + count], world)
        assert torch.equal(out, full), (rank, out)
        # cyclic tiles (blocks of 128 slots dealt round robin): gathered back into slot order
        Nc = 128 * world * 3
        slots = torch.from_numpy(mg.cyclic_slots(Nc, world, rank))
        vals = torch.arange(Nc, dtype=torch.float32) * 0.5
        assert torch.equal(mg.gather_cyclic(vals[slots], world), vals)
        owned = torch.zeros(Nc); owned[slots] = 1.0
        dist.all_reduce(owned); assert bool((owned == 1.0).all())         # the ranks' slots partition the swarm
        b0, c0 = mg.replicate_shard(10, world, rank)
        tot = torch.tensor([c0]); dist.all_reduce(tot); assert int(tot) == 10
        dist.destroy_process_group()
        print("OK", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2
