"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes, env-range sharding, action
slicing and the observation all-gather order (the physics itself is rank-local)."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from jiminy_b200.parallel import Shard, gather_observations
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    for n_total in (8, 9):
        sh = Shard(dist.get_rank(), dist.get_world_size(), n_total)
        lo, hi = sh.bounds
        actions = torch.arange(n_total * 3, dtype=torch.float64).reshape(n_total, 3)
        mine = sh.slice_actions(actions)
        assert mine.shape[0] == sh.n_local and float(mine[0, 0]) == 3.0 * lo
        # rank-local "physics": observation row = global env id
        obs = torch.stack([torch.full((5,), float(i), dtype=torch.float64) for i in range(lo, hi)])
        full = gather_observations(obs, sh)
        assert full.shape == (n_total, 5), full.shape
        assert torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float64)), full[:, 0]
    dist.barrier()
    dist.destroy_process_group()
    print("OK", flush=True)
""") % ROOT


def test_sharding_and_gather_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "OK" in o, o


def test_shard_bounds_cover_everything():
    from jiminy_b200.parallel import Shard
    for n, w in ((4096, 8), (4097, 8), (5, 4), (3, 4)):
        spans = [Shard(r, w, n).bounds for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
