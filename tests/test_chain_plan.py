"""Host logic of the backward chain launch (csrc/wc_small.cu::wc_bwd_chain), checked WITHOUT a GPU through the dry-run
hook wc_debug_bwd_chain_plan (include/wc_b200_debug.h): which jobs run in which stage, how the sums over classes and the
k range of the stage-2 product are cut into parts, and that a part never adds into a buffer the caller did not zero."""
import ctypes

import pytest

from wc_gan_b200._lib import lib

lib.wc_debug_bwd_chain_plan.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
lib.wc_debug_bwd_chain_plan.restype = ctypes.c_int

U, CC, F, ZRED, ZWS = 1, 2, 4, 8, 16       # `what` bits of the hook
SMM, VEC = 0, 1                           # job types (CJ_SMM, CJ_VEC)
EPI_PACK, EPI_PSI = 1, 2
SMS = 148


def plan(C, g, K, E, what):
    out = (ctypes.c_int * 512)()
    rc = lib.wc_debug_bwd_chain_plan(C, g, K, E, what, out, 512)
    assert rc == 0, rc
    ns, grid, ne = out[0], out[1], out[2]
    ents = [dict(type=out[9 + 6 * e], stage=out[10 + 6 * e], n=out[11 + 6 * e], l0=out[12 + 6 * e],
                 chunks_per=out[13 + 6 * e], atomic=out[14 + 6 * e] & 1, epi=out[14 + 6 * e] >> 1) for e in range(ne)]
    return dict(stages=ns, grid=grid, totals=list(out[3:3 + ns]), ents=ents)


def tiles(C, g):
    return (C // g) * (g // 32) ** 2


def test_headline_uconv():
    """cfg3: d + uconv, C = g = 256, K = 1: four stages, every stage one round; only the Y product of stage 2 is split."""
    p = plan(256, 256, 1, 0, U | ZRED | ZWS)
    assert p['stages'] == 4 and p['grid'] <= SMS and all(t <= SMS for t in p['totals'])
    smm = [e for e in p['ents'] if e['type'] == SMM]
    by_stage = {s: [e for e in smm if e['stage'] == s] for s in range(4)}
    assert [e['atomic'] for e in by_stage[0]] == [0, 0] and all(e['n'] == 64 and e['chunks_per'] == 4 for e in by_stage[0])
    psi, = by_stage[1]
    assert psi['epi'] == EPI_PSI and not psi['atomic'] and psi['n'] == 64
    y, = by_stage[2]
    assert y['atomic'] and y['n'] == 128 and y['chunks_per'] == 2
    a, = by_stage[3]
    assert a['epi'] == EPI_PACK and not a['atomic'] and a['chunks_per'] == 4


def test_nothing_is_split_without_zeroed_buffers():
    for what in (U, U | F, U | CC):
        for C, g, K in ((256, 256, 1), (128, 128, 10), (256, 64, 6)):
            p = plan(C, g, K, 4, what)
            assert not any(e['atomic'] for e in p['ents'])
            dweff = [e for e in p['ents'] if e['type'] == SMM and e['stage'] == 0][0]
            assert dweff['n'] == tiles(C, g) and dweff['chunks_per'] == K * ((C + 63) // 64)


@pytest.mark.parametrize('what', [U | F, U | CC, CC, F])
def test_cfg2_conditional_single_rounds(what):
    """cfg2 (C = g = 128, K = 10): the class sum is cut into 5 parts of 2 classes, Q fills the rest of stage 0 and runs on in
    stage 1, and every stage fits the grid once."""
    p = plan(128, 128, 10, 4, what | ZRED | ZWS)
    assert p['stages'] == 4 and all(t <= SMS for t in p['totals']), p['totals']
    s0 = [e for e in p['ents'] if e['type'] == SMM and e['stage'] == 0]
    dweff = s0[0]
    assert dweff['atomic'] and dweff['n'] == 16 * 5 and dweff['chunks_per'] == 4
    q = [e for e in p['ents'] if e['type'] == SMM and not e['atomic'] and e['epi'] == 0 and e['chunks_per'] == 2 and e['stage'] in (0, 1)]
    assert len(q) == 2 and q[0]['stage'] == 0 and q[1]['stage'] == 1
    assert q[0]['l0'] == 0 and q[1]['l0'] == q[0]['n'] and q[0]['n'] + q[1]['n'] == 10 * 16
    assert p['totals'][0] == SMS                                   # stage 0 is filled exactly
    # a 2-chunk tile is not cut further (a 1-chunk part costs more than it saves)
    assert not any(e['atomic'] for e in p['ents'] if e['stage'] >= 1)


def test_many_classes_factorised():
    """cfg5 (K = 200, E = 15): five stages (0a, 0b, Psi, Y, A); the sum over the 16 basis matrices is split."""
    p = plan(128, 128, 200, 15, U | F | ZRED | ZWS)
    assert p['stages'] == 5
    dweff = [e for e in p['ents'] if e['type'] == SMM and e['atomic']]
    assert len(dweff) == 1 and dweff[0]['stage'] == 1 and dweff[0]['n'] == 16 * 8 and dweff[0]['chunks_per'] == 4


@pytest.mark.parametrize('C,g', [(64, 64), (128, 128), (128, 32), (256, 256), (256, 64), (192, 96)])
@pytest.mark.parametrize('K,E,what', [(1, 0, U), (3, 2, U | F), (10, 4, U | CC), (16, 4, U | F), (40, 4, F), (7, 0, CC)])
def test_plan_invariants(C, g, K, E, what):
    for zero in (0, ZRED, ZRED | ZWS):
        p = plan(C, g, K, E, what | zero)
        nkc = (C + 63) // 64
        assert 4 <= p['stages'] <= 5 and 1 <= p['grid'] <= SMS
        assert all(e['n'] > 0 for e in p['ents'])
        for s in range(p['stages']):
            assert p['totals'][s] == sum(e['n'] for e in p['ents'] if e['stage'] == s)
        for e in p['ents']:
            if e['atomic']:
                assert zero & ZRED if e['stage'] <= (1 if p['stages'] == 5 else 0) else zero & ZWS
                assert e['epi'] != EPI_PACK and e['chunks_per'] >= 1
        # the critical products are all there: Psi once, the operand image of A in the last stage when g == C
        assert sum(1 for e in p['ents'] if e['type'] == SMM and e['epi'] == EPI_PSI) == 1
        last = [e for e in p['ents'] if e['type'] == SMM and e['stage'] == p['stages'] - 1]
        assert len(last) == 1 and last[0]['n'] == tiles(C, g) and last[0]['chunks_per'] == (g + 63) // 64
        assert (last[0]['epi'] == EPI_PACK) == (g == C)
