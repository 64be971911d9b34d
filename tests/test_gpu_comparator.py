"""The reference's OWN GPU path as a second checker: its GPU task function and kernels (K1 DMMA, K2 FMA),
compiled unmodified for sm_100a (oracle/_ref/libccsdt_refgpu_*.so; oracle/Makefile `refgpu`), run on the
same tensors as this repo's fused kernel.  Three independent evaluations (reference CPU fixture, reference
GPU kernels, ours) must agree to 1e-9 Eh.  /root/reference is not needed at run time: the prebuilt
libraries travel with the snapshot; without them the tests skip."""
import os

import numpy as np
import pytest

from exachem_b200 import driver as drv, synthetic as syn

pytestmark = pytest.mark.gpu
ATOL = 1e-9


def _ref_gpu(kind):
    from oracle.oracle import REF_GPU_SO, ReferenceGPU
    if not os.path.exists(REF_GPU_SO[kind]):
        pytest.skip("oracle/_ref GPU comparator was not built (needs /root/reference at build time)")
    return ReferenceGPU(kind)


@pytest.mark.parametrize("kind", ["tc", "fma"])
def test_reference_gpu_kernels_agree_with_ours_and_with_the_cpu_fixture(kind):
    from oracle.oracle import Oracle
    ref = _ref_gpu(kind)
    orc = Oracle()
    for oa, ob, va, vb, ts, seed in ((4, 4, 10, 10, 6, 2024), (5, 5, 19, 19, 28, 7), (3, 3, 9, 9, 4, 11)):
        osp = orc.tiles(oa, ob, va, vb, ts)
        sp = drv.setup_mo_space(oa, ob, va, vb, ts)
        T = syn.dense_all(syn.Orbitals(oa, ob, va, vb), seed)
        out, trace = ref.execute(osp, T, True, tilesize=ts)
        ctx = drv.Context(0)
        try:
            ctx.set_space(sp, T["evl"], True)
            for tid, k in ((drv.T1, "t1"), (drv.T2, "t2"), (drv.V_IJAB, "v2ijab"), (drv.V_IJKA, "v2ijka"),
                           (drv.V_IABC, "v2iabc")):
                ctx.put_dense(tid, T[k])
            tasks, fac, _ = drv.enumerate_tasks(sp, True)
            e1, e2, _, pt = ctx.run(per_task_n=len(tasks))
        finally:
            ctx.close()
        assert abs(out[0] - e1) <= ATOL and abs(out[1] - e2) <= ATOL, (kind, out[:2], e1, e2)
        # per task: kernel partial sums x factor
        assert len(trace) == len(tasks)
        np.testing.assert_allclose(trace[:, 8:10] * fac[:, None], pt, rtol=0, atol=ATOL)
        # extents the reference launcher was called with = the canonical task list's tiles
        np.testing.assert_array_equal(trace[:, :6].astype(np.int64), sp.k_range[tasks[:, :6]])
        # and the CPU-side oracle (pinned to the reference CPU kernel)
        r1, r2 = orc.run(osp, T, True)
        assert abs(out[0] - r1) <= ATOL and abs(out[1] - r2) <= ATOL
    ref.release()
