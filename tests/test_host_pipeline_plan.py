"""The schedule of the slab-pipelined host step (mw_host_pipeline_plan, csrc/dycore.cu) is pure host logic: checked here
without a GPU.  For many grid sizes / tracer counts / sub-cycle counts the plan must
  * cover every row of every level exactly once,
  * only run an operation after everything it reads exists: the rows of the previous level within its reach (3 rows for
    a stage kernel, 1 row for the tracer finish after its stage kernel, 0 for the conversions), periodic in y,
  * only start a main-phase operation on rows whose upload it has waited for,
  * never let a stage overwrite, in place, rows that an earlier level still has to read (q[0] is both the RK register the
    last stage writes and the first stage's input; the flux / FCT scratch is shared by the three stages)."""
import ctypes as C
import numpy as np
import pytest

import miniweatherml_b200 as mw

HALO = 3


def plan(ny, rows, T, ncyc):
    L = mw.lib()
    buf = (C.c_int * (6 * 4096))()
    n, S = C.c_int(), C.c_int()
    L.mw_host_pipeline_plan.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    assert L.mw_host_pipeline_plan(ny, rows, T, ncyc, buf, 4096, C.byref(n), C.byref(S)) == 0
    ops = np.array(buf[:6 * n.value]).reshape(-1, 6)
    return ops, S.value


def rows_of(r0, r1, ny):
    return np.arange(r0, r1) % ny


@pytest.mark.parametrize("ny,rows,T,ncyc", [(512, 32, 1, 1), (512, 32, 0, 1), (1024, 32, 3, 1), (160, 8, 1, 1), (147, 8, 3, 1),
                                            (217, 16, 1, 1), (100, 16, 0, 1), (344, 24, 2, 3), (77, 8, 2, 1), (513, 40, 1, 2)])
def test_plan_covers_every_row_once_and_respects_dependencies(ny, rows, T, ncyc):
    ops, S = plan(ny, rows, T, ncyc)
    assert len(ops) > 0 and S == max(1, ny // rows)
    L = ops[:, 0].max() + 1
    assert L == 2 + 3 * ncyc * (2 if T else 1)
    kind = {int(l): int(k) for l, k in ops[:, :2]}
    done = np.zeros((L, ny), dtype=bool)
    done_at = np.full((L, ny), -1)                   # position in the plan at which (level, row) was produced
    uploaded = lambda s: ny if s == S - 1 else (s + 1) * rows
    seam_started = False
    for pos, (l, k, st, r0, r1, up) in enumerate(ops):
        rr = rows_of(r0, r1, ny)
        assert r0 >= 0 and r1 > r0 and r1 - r0 <= ny
        if k == 1:                                   # stage kernels work on whole tile rows
            assert r0 % 8 == 0 and (r1 % 8 == 0 or r1 == ny or r1 > ny)
            if r1 > ny:
                assert (r1 - ny) % 8 == 0
        assert not done[l, rr].any(), "rows done twice"
        if up >= 0:
            assert not seam_started
        else:
            seam_started = True
        if l == 0:
            assert up >= 0 and r1 <= uploaded(up)
        else:
            reach = HALO if k == 1 else (1 if k == 2 else 0)
            need = rows_of(r0 - reach, r1 + reach, ny) if r1 - r0 + 2 * reach <= ny else np.arange(ny)
            assert done[l - 1, need].all(), (pos, l, k, r0, r1)
            if k == 2:                               # the tracer finish also reads the state of the level before its stage kernel? no: own rows of the stage output
                assert done[l - 1, rr].all()
        done[l, rr] = True
        done_at[l, rr] = pos
    assert done.all()

    # in-place hazards.  Level numbering: 0 = c2d; per cycle and stage s: stage kernel (+ tracer finish).
    levels = sorted(kind)
    stage_levels = [l for l in levels if kind[l] == 1]
    for idx, l in enumerate(stage_levels):
        st = idx % 3
        # (a) the flux / FCT scratch written by stage kernel l is read by the previous stage's tracer finish (if any):
        #     rows within 1 of a row must have been finished by that tracer finish before the row is overwritten
        if T and idx > 0:
            prev_tr = stage_levels[idx - 1] + 1
            for r in range(ny):
                for d in (-1, 0, 1):
                    assert done_at[prev_tr, (r + d) % ny] < done_at[l, r]
        # (b) stage 3 (st == 2) writes q[0] in place: the first stage kernel of this cycle read q[0] rows within 3
        if st == 2:
            first = stage_levels[idx - 2]
            for r in range(0, ny, 1):
                for d in range(-HALO, HALO + 1):
                    assert done_at[first, (r + d) % ny] < done_at[l, r]
            # and the second stage (reads q0 = q[0] at its own rows)
            second = stage_levels[idx - 1]
            assert (done_at[second] < done_at[l]).all()
        # (c) with sub-cycling the next cycle's stage kernel overwrites q[1] / q[2] that the previous cycle's next stage read
        if idx >= 3:
            reader = stage_levels[idx - 2]          # stage (st+1) of the previous cycle read this buffer with halo
            for r in range(ny):
                for d in range(-HALO, HALO + 1):
                    assert done_at[reader, (r + d) % ny] < done_at[l, r]


def test_plan_refuses_grids_that_are_too_small():
    ops, S = plan(40, 8, 3, 1)
    assert len(ops) == 0                             # the caller falls back to the serial path


def test_plan_lag_is_small():
    """the download of a slab can start well before the upload ends: the last level trails the upload by 25 rows (24 without tracers)"""
    ops, S = plan(512, 32, 1, 1)
    L = ops[:, 0].max() + 1
    first_d2c = [o for o in ops if o[0] == L - 1][0]
    assert first_d2c[5] == 1 and first_d2c[3] == 25  # after the second slab upload, from row 25
    ops0, _ = plan(512, 32, 0, 1)
    assert [o for o in ops0 if o[0] == ops0[:, 0].max()][0][3] == 24
