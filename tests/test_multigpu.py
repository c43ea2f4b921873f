"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): two ranks in ONE process
(upsp_gpu_connect_local), frames sharded by apportion(), node-major rows stored straight into
the peer's buffer, rank-ordered sum/sum-sq over peer memory.  Compared with the oracle run with
two simulated MPI ranks."""
import numpy as np
import pytest

from chain import Case, cp_errors, monomial_mass, push_all, run_oracle, same_bits, setup_ctx

pytestmark = pytest.mark.gpu


def _run_ranks(up, orc, case, keep_frame_major, same_device=False, R=2, **kw):
    ctxs, slices = [], []
    for r in range(R):
        g, sl = setup_ctx(up, orc, case, rank=r, n_ranks=R, device=0 if same_device else r,
                          keep_frame_major=keep_frame_major, **kw)
        ctxs.append(g)
        slices.append(sl)
    up.connect_local(ctxs)
    for g, sl in zip(ctxs, slices):
        push_all(up, orc, g, case, sl)
    for g in ctxs:
        g.sync()                      # == MPI_Barrier before the reduce
    for g in ctxs:
        g.finish_phase1()
    for g in ctxs:
        g.transpose()
    for g in ctxs:
        g.sync()                      # == MPI_Barrier after the transpose
    out = dict(itrans=[], ptrans=[], rms2=[], avg2=[], gain=[])
    stats = [g.read_phase1_stats() for g in ctxs]
    for g in ctxs:
        out["itrans"].append(g.read_intensity_transpose())
        g.phase2(case.cal, case.qbar, case.ps, case.steady, case.temp, case.degree)
        out["ptrans"].append(g.read_pressure_transpose())
        r2, a2, g2 = g.read_phase2_stats()
        out["rms2"].append(r2)
        out["avg2"].append(a2)
        out["gain"].append(g2)
    res = {k: np.concatenate(v, 0) for k, v in out.items()}
    res["avg"], res["rms"], res["coverage"] = stats[0]
    # every rank must hold bit-identical phase-1 statistics (rank-ordered reduction)
    for r in range(1, R):
        for k in range(3):
            assert same_bits(stats[0][k], stats[r][k])
    for g in ctxs:
        g.close()
    return res


@pytest.mark.parametrize("keep_frame_major", [False, True], ids=["fused", "frame-major+transpose"])
@pytest.mark.parametrize("same_device", [False, True], ids=["2gpus", "2ranks-1gpu"])
def test_two_ranks_match_oracle(up, orc, gpu, keep_frame_major, same_device):
    if not same_device and gpu < 2:
        pytest.skip("needs 2 GPUs")
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=75, n_nodes=3001, registration=True, patches=True, overlap=True,
                seed=21, fmt="p12")
    ref = run_oracle(orc, case, n_ranks=2)
    got = _run_ranks(up, orc, case, keep_frame_major, same_device)
    assert same_bits(got["avg"], ref["avg"]) and same_bits(got["rms"], ref["rms"])
    assert same_bits(got["itrans"], ref["itrans"]), "intensity_transpose not bit-exact across ranks"
    assert same_bits(got["gain"], ref["gain"])
    exact = run_oracle(orc, case, n_ranks=2, exact_fit=True)
    e_exact, _ = cp_errors(case, exact, got)
    cond = 8 * np.finfo(np.float32).eps * monomial_mass(orc, ref)
    assert np.all(e_exact <= 1e-6 + cond)
    e_op, _ = cp_errors(case, ref, got)
    noise, _ = cp_errors(case, ref, exact)
    assert np.all(e_op <= 1e-5 + noise + cond)


@pytest.mark.parametrize("R,staged,ship", [(3, None, None), (4, "1", None), (8, "3", None), (8, None, None),
                                           (4, "3", "sm"), (8, "7", "sm"), (3, "1", "sm")])
def test_many_ranks_one_gpu(up, orc, gpu, R, staged, ship, monkeypatch):
    """R ranks as R contexts on one device.  >= 3 ranks store rows straight into the owners' buffers
    in 128-byte segments (default); UPSP_STAGED_PEERS=k sends the next k ranks' rows through the
    staging block + copy engines instead (mixed exchange).  Small batches so that both staging
    buffers and several exchange rounds are used."""
    import upsp_b200
    if staged is not None:
        monkeypatch.setenv("UPSP_STAGED_PEERS", staged)
    if ship is not None:          # staged rows shipped by the SM kernel k_ship_rows instead of the copy engines
        monkeypatch.setenv("UPSP_SHIP", ship)
    case = Case(upsp_b200.synth, n_frames=96, n_nodes=2003, registration=True, patches=True, overlap=True,
                seed=33, fmt="p12")
    ref = run_oracle(orc, case, n_ranks=R)
    got = _run_ranks(up, orc, case, False, True, R=R, batch_frames=8)
    assert same_bits(got["avg"], ref["avg"]) and same_bits(got["rms"], ref["rms"])
    assert same_bits(got["itrans"], ref["itrans"]), "intensity_transpose not bit-exact across ranks"
    assert same_bits(got["gain"], ref["gain"])


@pytest.mark.parametrize("R", [2, 4])
def test_ranks_batch_blocked_rows(up, orc, gpu, R):
    """Several ranks, every rank the same multiple-of-8 number of frames, power-of-two batch: the 16-bit rows of the plain
    nodes live in the batch-blocked layout [source rank * batches + local batch][node][batch length] (what a projection
    batch stores into a peer is one contiguous region), phase 2 and the readers follow it.  528 frames = 2 x 264 or
    4 x 132: the last block of every rank is partly filled."""
    import upsp_b200
    case = Case(upsp_b200.synth, n_frames=528, n_nodes=2003, registration=True, patches=True, seed=35, fmt="p12")
    ref = run_oracle(orc, case, n_ranks=R)
    got = _run_ranks(up, orc, case, False, True, R=R, batch_frames=64)
    assert same_bits(got["avg"], ref["avg"]) and same_bits(got["rms"], ref["rms"])
    assert same_bits(got["itrans"], ref["itrans"]), "intensity_transpose not bit-exact across ranks"
    assert same_bits(got["gain"], ref["gain"])
    exact = run_oracle(orc, case, n_ranks=R, exact_fit=True)
    e_exact, _ = cp_errors(case, exact, got)
    cond = 8 * np.finfo(np.float32).eps * monomial_mass(orc, ref)
    assert np.all(e_exact <= 1e-6 + cond)
