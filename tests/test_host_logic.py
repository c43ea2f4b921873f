"""CPU tests of the host-side logic (no GPU): apportion slices as the library computes them,
and the N>1 plumbing bench.py uses (handle all-gather + barriers) over gloo, world_size 2."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT


def test_oracle_apportion_matches_reference_rule(orc):
    # psp_process.cpp:611-624: first `remainder` bins get one extra element
    for value, bins in ((10, 3), (20000, 8), (7, 8), (0, 4), (500000, 7)):
        s, e = orc.apportion(value, bins)
        assert e.sum() == value
        assert list(e) == [value // bins + (b < value % bins) for b in range(bins)]
        assert list(s) == list(np.concatenate([[0], np.cumsum(e)[:-1]]))


def test_two_rank_oracle_equals_one_rank(orc):
    """The rank-sharded reference flow (phase 1 per rank, reduce, global transpose, phase 2 per
    node slice) gives the same bits as the single-rank flow."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import upsp_b200
    from chain import Case, run_oracle, same_bits
    case = Case(upsp_b200.synth, n_frames=21, n_nodes=500, registration=True, patches=True, overlap=True, seed=5)
    a = run_oracle(orc, case, n_ranks=1)
    b = run_oracle(orc, case, n_ranks=3)
    for k in ("intensity", "itrans", "avg", "rms", "coverage", "ptrans", "gain"):
        assert same_bits(a[k], b[k]), k


def test_gloo_world2_handle_exchange_and_barriers(tmp_path):
    """bench.py's multi-rank plumbing: all_gather of 64-byte handles + barriers + MAX reduce,
    torch.distributed gloo backend, 2 processes on this CPU box."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import bench
        dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
        rank = dist.get_rank()
        h = torch.frombuffer(bytearray(bytes([rank + 1]) * 64), dtype=torch.uint8).clone()
        allh = [torch.empty_like(h) for _ in range(2)]
        dist.all_gather(allh, h)
        blob = b"".join(bytes(x.numpy().tobytes()) for x in allh)
        assert blob == bytes([1]) * 64 + bytes([2]) * 64
        bench.barrier(dist)
        assert bench.allmax(dist, float(rank + 1)) == 2.0
        print("ok", rank)
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
