"""Index model of the batch-blocked 16-bit row layout of multi-rank runs (csrc/upsp_gpu.cu
ensure_proj_mode / process_batch_impl, csrc/kernels_project_tma.cuh row pointer, csrc/kernels_phase2.cuh p2_off16,
csrc/upsp_gpu.cu k_itrans_rows_f32).  The destination rank d keeps [R * K][N_d][L] 16-bit integers: block s*K + j holds frames
[s*F_loc + j*L, s*F_loc + (j+1)*L) of source rank s for each of its N_d nodes (the reference's intensity_transpose row of a
node, cpp/exec/psp_process.cpp:719-767, is the concatenation of those pieces in frame order).  Three pieces of arithmetic
must agree: where the projection of rank s stores frame `off + i` of a batch, where phase 2 looks for global frame f (a
multiply-high division by F_loc with one correction), and where the ABI's reader looks.  No GPU: integers only."""
import numpy as np
import pytest


def writer_offset(s, off, i, n_local_index, N_d, L, K):
    """Element offset in rank d's buffer of local frame off + i of source rank s, node n (batches never straddle a block)."""
    blk_index = s * K + off // L
    j0 = off % L
    return (blk_index * N_d + n_local_index) * L + j0 + i


def p2_off16(f, li, N_d, L, K, F_loc):
    """kernels_phase2.cuh p2_off16: magic = floor(2^32 / F_loc) + 1, r = umulhi(f, magic), one step too high at most."""
    magic = ((1 << 32) // F_loc + 1) & 0xFFFFFFFF
    r = (f * magic) >> 32
    if r * F_loc > f:
        r -= 1
    o = f - r * F_loc
    lg = L.bit_length() - 1
    blk = r * K + (o >> lg)
    return ((blk * N_d + li) << lg) + (o & (L - 1))


def reader_offset(f, row, N_d, L, K, F_loc):
    """upsp_gpu.cu k_itrans_rows_f32, blocked branch (plain division)."""
    src, o = divmod(f, F_loc)
    return ((src * K + o // L) * N_d + row) * L + o % L


@pytest.mark.parametrize("R,F_loc,L", [(2, 264, 64), (4, 132 // 4 * 8, 64), (8, 20000, 256), (2, 20000, 256), (3, 8, 64), (5, 4096, 128)])
def test_writer_phase2_and_reader_agree(R, F_loc, L):
    assert F_loc % 8 == 0 and L & (L - 1) == 0
    K = -(-F_loc // L)
    N_d = 7
    F = R * F_loc
    seen = {}
    for s in range(R):
        off = 0
        while off < F_loc:                                   # upsp_gpu_process_frames: batches are cut at block edges
            nb = min(L - off % L, F_loc - off)
            for i in range(0, nb, max(1, nb // 5)):          # sample the batch (its first frame included)
                for n in (0, N_d - 1):
                    w = writer_offset(s, off, i, n, N_d, L, K)
                    f = s * F_loc + off + i
                    assert w == p2_off16(f, n, N_d, L, K, F_loc) == reader_offset(f, n, N_d, L, K, F_loc)
                    assert 0 <= w < R * K * N_d * L
                    assert seen.setdefault(w, (f, n)) == (f, n)      # no two node-frames share an element
            off += nb
    # quads: 4 consecutive frames from a multiple of 4 never straddle a block or a rank (F_loc % 8 == 0, L % 8 == 0)
    for f in range(0, F, 4):
        a = p2_off16(f, 3, N_d, L, K, F_loc)
        assert [p2_off16(f + j, 3, N_d, L, K, F_loc) for j in range(4)] == [a, a + 1, a + 2, a + 3]


def test_magic_division_is_exact_for_every_frame():
    for F_loc in (8, 24, 264, 1000, 4096, 20000, 65528, 1 << 20):
        magic = ((1 << 32) // F_loc + 1) & 0xFFFFFFFF
        f = np.arange(0, min(16 * F_loc, 1 << 24), dtype=np.uint64)
        r = (f * np.uint64(magic)) >> np.uint64(32)
        r = r - (r * np.uint64(F_loc) > f)
        assert np.array_equal(r, f // np.uint64(F_loc))
