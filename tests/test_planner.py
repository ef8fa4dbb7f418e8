"""Host logic of the tcgen05 launch planners (`conv_tc.cu` / `conv_tf.cu` plan()), through `sb200_debug_plan`: planning
only, nothing is launched, so this runs without a GPU.

The property that matters most: the reference's `speak_batch` is a loop of B=1 runs, and this implementation promises
the same bits for an utterance whether it is synthesised alone or inside a batch (`test_batched_equals_sequential`,
`test_frontends_on_the_device` check it on the device).  The planners pick tile widths, ring depths and tile pairing from
the launch SIZE -- none of which changes a summation order -- but the choices that DO change arithmetic (cat mode: two
MMAs per step instead of three; the TMA-staged epilogue's rounding; the chunk length of the flushed accumulation) must
depend on the layer's shape alone."""
import ctypes as C

import pytest

from sonata_b200 import _native as N
from sonata_b200 import voicegen

ROWS = (100, 128, 700, 3000, 20_000, 300_000, 4_000_000)
SMEM_MAX = 227 * 1024
SMS = 148                                            # the planners assume a B200 when no device is visible
ACT_NONE, ACT_RELU, ACT_GATE = 0, 1, 2


def plan(backend, rows, cin, cout, k, dil, act=ACT_NONE, res=0, acc=0):
    o = (C.c_int32 * 16)()
    rc = N.lib().sb200_debug_plan(backend, rows, cin, cout, k, dil, act, res, acc, o)
    return None if rc else list(o)


def decoder_and_flow_layers(q):
    """(cin, cout, k, dil, act, res, acc) of every conv_tc layer of a voice (the phase-fused transposed convs aside)."""
    a = voicegen.ARCH[q]
    H, half = a["hidden"], a["inter"] // 2
    L = {(half, H, 1, 1, ACT_NONE, 0, 0), (H, 2 * H, a["flow_kernel"], 1, ACT_GATE, 0, 0), (H, 2 * H, 1, 1, ACT_NONE, 0, 1),
         (H, H, 1, 1, ACT_NONE, 0, 1), (H, half, 1, 1, ACT_NONE, 0, 0), (a["inter"], a["up_init"], 7, 1, ACT_NONE, 0, 0)}
    ch = a["up_init"]
    for _ in a["up_rates"]:
        ch //= 2
        for k, dils in zip(a["res_kernels"], a["res_dils"]):
            for d in dils:
                if a["resblock"] == 2:
                    L |= {(ch, ch, k, d, ACT_NONE, 1, acc) for acc in (0, 1)}
                else:
                    L |= {(ch, ch, k, d, ACT_NONE, 0, 0)} | {(ch, ch, k, 1, ACT_NONE, 1, acc) for acc in (0, 1)}
    return sorted(L)


def encoder_layers(q):
    a = voicegen.ARCH[q]
    H, F, k = a["hidden"], a["filter"], a["kernel"]
    return [(H, 3 * H, 1, 1, ACT_NONE, 0, 0), (H, H, 1, 1, ACT_NONE, 0, 0), (H, F, k, 1, ACT_RELU, 0, 0), (F, H, k, 1, ACT_NONE, 0, 0),
            (H, 2 * a["inter"], 1, 1, ACT_NONE, 0, 0)]


@pytest.mark.parametrize("q", ["medium", "high"])
def test_arithmetic_class_does_not_depend_on_launch_size(q):
    for lay in decoder_and_flow_layers(q):
        seen = set()
        for rows in ROWS:
            p = plan(1, rows, *lay)
            assert p is not None, (lay, rows)
            nt, wnt, mt, ntn, resident, cat, tma_epi, pairs, na, ws, nstg, smem = p[:12]
            seen.add((cat, tma_epi))
            assert smem <= SMEM_MAX and na >= 2 and wnt % nt == 0 and nt % 32 == 0, (lay, rows, p)
            if nt < wnt:                       # narrow tiles: only while no SM would get a second tile, never with the
                assert mt * ntn <= SMS and not tma_epi and not cat, (lay, rows, p)      # TMA epilogue / cat arithmetic
            if pairs:                          # shared weight ring: streamed weights on launches worth two tiles per SM
                assert not resident and mt * ntn >= 2 * SMS, (lay, rows, p)
            if cat:
                assert tma_epi and nt <= 64, (lay, rows, p)
        assert len(seen) == 1, (lay, seen)
    for lay in encoder_layers(q):
        chunks = set()
        for rows in ROWS:
            p = plan(2, rows, *lay)
            assert p is not None, (lay, rows)
            nth, wnth, mp, ntn, na, nw, chunk_kb, smem = p[:8]
            chunks.add(chunk_kb)
            assert smem <= SMEM_MAX and wnth % nth == 0 and nth in (32, 64, 96), (lay, rows, p)
            if nth < wnth:
                assert mp * (lay[1] // 32) <= SMS, (lay, rows, p)
        assert len(chunks) == 1, (lay, chunks)


def test_hot_layers_get_the_configurations_design_md_describes():
    # 32-channel ResBlock conv (dec.mrf2 at C2 size): resident weights, cat mode, TMA-staged epilogue, four stages
    p = plan(1, 14_700_000, 32, 32, 3, 1, ACT_NONE, 1, 0)
    assert p[4:7] == [1, 1, 1] and p[8] == 4
    # 64-channel k3: the same epilogue through two 32-column chunks of one 64-column tile
    p = plan(1, 3_700_000, 64, 64, 3, 1, ACT_NONE, 1, 0)
    assert p[0] == 64 and p[4] == 1 and p[6] == 1
    # 128 channels: row-per-thread epilogue; flow in_layer: streamed weights, tile pairs on a big launch
    assert plan(1, 460_000, 128, 128, 3, 1, ACT_NONE, 1, 0)[6] == 0
    p = plan(1, 57_600, 192, 384, 5, 1, ACT_GATE, 0, 0)
    assert p[0] == 128 and p[4] == 0 and p[7] == 1
    # the same layer for one utterance: 32-column parts of the 128-row images, one tile per CTA
    p = plan(1, 900, 192, 384, 5, 1, ACT_GATE, 0, 0)
    assert (p[0], p[1], p[7]) == (32, 128, 0)
    # unsupported shapes are refused, not mis-planned
    assert plan(1, 1000, 48, 64, 3, 1) is None and plan(2, 1000, 192, 100, 1, 1) is None
