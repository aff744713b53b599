"""Drop-in test of the JM boundary (capture plan (i)): the stock JM 18.5 encoder and the same
encoder with me_fullsearch.o replaced by integration/jm/b2me_jm_shim.c + libb2me.so must write
byte-identical bitstreams and reconstructions on BASELINE config 1 (synthetic QCIF, IPPP, full search
+-16, 1 reference, QP 28, SAD integer + SATD sub-pel).  Both binaries are prebuilt by oracle/Makefile.jm
in the build container (oracle/_ref/ travels to the GPU box)."""
import os
import tempfile

import pytest

from h264_b200 import synth
from oracle import jm_run

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
have = all(os.path.exists(os.path.join(REF, f)) for f in ("lencod", "lencod_b2", "encoder.cfg"))


def _encode(exe, yuv, W, H, frames, out, **kw):
    log = jm_run.run_lencod(yuv, W, H, frames, out, exe=exe, **kw)
    return open(os.path.join(out, "out.264"), "rb").read(), open(os.path.join(out, "rec.yuv"), "rb").read(), log


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref/lencod{,_b2} not built (needs /root/reference at build time)")
@pytest.mark.parametrize("frames,nrefs,sr,qp", [(30, 1, 16, 28), (6, 2, 8, 36)])
def test_lencod_with_cuda_motion_search_is_bit_identical(frames, nrefs, sr, qp):
    W, H = 176, 144
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=20261018))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=nrefs, search_range=sr, qp=qp)
        b = _encode("lencod_b2", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=nrefs, search_range=sr, qp=qp,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref/lencod{,_b2} not built (needs /root/reference at build time)")
def test_lencod_b_slices_with_cuda_bipred_search_are_bit_identical():
    """IBPBP with BiPredMotionEstimation: full_search_bipred_motion_estimation / sub_pel_bipred_motion_estimation (and the
    list-1 single-list searches) served by libb2me.so; 3 refinement iterations, dual-level sub-pel."""
    W, H, frames = 176, 144, 7
    extra = ("NumberBFrames=1", "BiPredMotionEstimation=1", "BiPredMERefinements=3", "BiPredMESearchRange=8", "BiPredMESubPel=2",
             "HierarchicalCoding=0", "BReferencePictures=0", "QPBSlice=30", "DirectModeType=1", "BList0References=0",
             "BList1References=1", "BiPredSearch16x16=1", "BiPredSearch16x8=1", "BiPredSearch8x16=1", "BiPredSearch8x8=0")
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=77))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=2, search_range=8, qp=30, extra=extra)
        b = _encode("lencod_b2", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=2, search_range=8, qp=30, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) bi-predictive calls", b[2])
        assert m and int(m.group(1)) > 1000, b[2][-400:]      # the B slices really went through b2me_bipred_search


@pytest.mark.gpu
@pytest.mark.skipif(not have, reason="oracle/_ref/lencod{,_b2} not built (needs /root/reference at build time)")
@pytest.mark.parametrize("mode,extra", [
    # EPZS keeps its integer search in JM; its sub-pel stages are the 81-position functions of me_fullsearch.o (EPZSSubPelME = 2)
    (3, ("EPZSSubPelME=2", "EPZSSubPelMEBiPred=2", "EPZSSubPelGrid=0")),
    # fast full search (me_fullfast.o, JM) for the single lists; bi-predictive search and every sub-pel stage from libb2me.so
    (0, ()),
])
def test_other_search_modes_share_the_cuda_subpel_and_bipred_functions(mode, extra):
    W, H, frames = 176, 144, 5
    extra = tuple(extra) + ("NumberBFrames=1", "BiPredMotionEstimation=1", "BiPredMERefinements=1", "BiPredMESearchRange=8", "BiPredMESubPel=2",
                            "HierarchicalCoding=0", "BReferencePictures=0", "QPBSlice=30", "BList1References=1")
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=5))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=2, search_range=8, qp=30, search_mode=mode, extra=extra)
        b = _encode("lencod_b2", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=2, search_range=8, qp=30, search_mode=mode, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) sub-pel refinements, (\d+) bi-predictive calls", b[2])
        assert m and int(m.group(1)) > 1000 and int(m.group(2)) > 100, b[2][-400:]


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(os.path.join(REF, "lencod_b2d"))), reason="oracle/_ref/lencod_b2d not built")
def test_epzs_with_every_distortion_on_the_gpu_is_bit_identical():
    """lencod_b2d: computeSAD / computeSATD of me_distortion.o served by b2me_distortion_candidates (one call per candidate).
    EPZS (SearchMode 3, its own sub-pel search) keeps all its control flow in JM and takes every distortion from the GPU."""
    W, H, frames = 176, 144, 3
    extra = ("EPZSSubPelME=1", "EPZSSubPelGrid=0")
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=9))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=1, search_range=16, qp=28, search_mode=3, extra=extra)
        b = _encode("lencod_b2d", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=1, search_range=16, qp=28, search_mode=3, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) distortion calls", b[2])
        assert m and int(m.group(1)) > 10000, b[2][-400:]


def _fading_clip(W, H, frames, seed):
    """a clip whose luma fades from frame to frame, so that explicit weighted prediction finds weights other than the default"""
    import numpy as np
    raw = np.frombuffer(synth.yuv420_sequence(W, H, frames, seed=seed), np.uint8).reshape(frames, W * H * 3 // 2).copy()
    for i in range(frames):
        y = raw[i, :W * H].astype(np.float64)
        raw[i, :W * H] = np.clip(y * (1.0 - 0.09 * i) + 3 * i, 0, 255).astype(np.uint8)
    return raw.tobytes()


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(os.path.join(REF, "lencod_b2d"))), reason="oracle/_ref/lencod_b2d not built")
@pytest.mark.parametrize("exe,mode,extra,counter", [
    # B slices under EPZS with its own bi-predictive search: computeBiPredSAD1 / SATD1 per candidate pair, the block distortions of BIDPartitionCost
    ("lencod_b2d", 3, ("EPZSSubPelME=1", "EPZSSubPelMEBiPred=1", "EPZSSubPelGrid=0", "NumberBFrames=1", "BiPredMotionEstimation=1", "HierarchicalCoding=0",
                       "BReferencePictures=0", "QPBSlice=30", "BList1References=1", "BiPredMESearchRange=8", "BiPredMERefinements=1"), r"(\d+) bi-predictive distortion calls"),
    # explicit weighted prediction on a fading clip, weighted reference ME: computeSADWP / SATDWP through weighted planes (EPZS)
    ("lencod_b2d", 3, ("EPZSSubPelME=1", "EPZSSubPelGrid=0", "WeightedPrediction=1", "UseWeightedReferenceME=1"), r"(\d+) distortion calls"),
    # the same configuration under the full search: full_search / sub_pel of me_fullsearch.o on weighted planes
    ("lencod_b2", -1, ("WeightedPrediction=1", "UseWeightedReferenceME=1"), r"(\d+) integer searches"),
])
def test_weighted_and_bipredictive_distortions_on_the_gpu_are_bit_identical(exe, mode, extra, counter):
    """The rest of me_distortion.o in the drop-in (lencod_b2d: that object is out of the link, all 22 symbols come from the shim) and
    weighted single-list motion estimation of the full search (b2me_set_ref_weights)."""
    W, H, frames = 176, 144, 4
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(_fading_clip(W, H, frames, 23))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=2, search_range=8, qp=30, search_mode=mode, extra=extra)
        b = _encode(exe, yuv, W, H, frames, os.path.join(d, "b2"), nrefs=2, search_range=8, qp=30, search_mode=mode, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(counter, b[2])
        assert m and int(m.group(1)) > 500, b[2][-500:]


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(os.path.join(REF, "lencod_b2e"))), reason="oracle/_ref/lencod_b2e not built")
@pytest.mark.parametrize("frames,nrefs,sr,extra", [
    (6, 2, 16, ("EPZSSubPelME=2", "EPZSSubPelGrid=0")),                                              # P slices, two references (the prevSad exits of ref > 0)
    (4, 1, 32, ("EPZSSubPelME=2", "EPZSSubPelGrid=0", "EPZSPattern=5", "EPZSDualRefinement=6")),     # PMVFAST pattern chain, dual refinement with the large diamond
    (5, 2, 8, ("EPZSSubPelME=2", "EPZSSubPelMEBiPred=2", "EPZSSubPelGrid=0", "NumberBFrames=1", "HierarchicalCoding=0", "BReferencePictures=0",
               "QPBSlice=30", "BList1References=1", "EPZSFixedPredictors=3", "EPZSTemporal=1", "EPZSSpatialMem=1", "EPZSBlockType=1")),   # B slices: list 1, temporal + memory predictors
])
def test_epzs_integer_search_on_the_gpu_is_bit_identical(frames, nrefs, sr, extra):
    """lencod_b2e: EPZS_motion_estimation / EPZS_subMB_motion_estimation (JM/lencod/src/me_epzs.c:54-407, :417-750)
    served by b2me_epzs_search -- median check, predictor scan, pattern refinement and the dual round on the device, one call per
    block; predictors, thresholds and pattern tables are the reference's own (me_epzs_common.o)."""
    W, H = 176, 144
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=31))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=nrefs, search_range=sr, qp=28, search_mode=3, extra=extra)
        b = _encode("lencod_b2e", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=nrefs, search_range=sr, qp=28, search_mode=3, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) EPZS searches \((\d+) search points\)", b[2])
        assert m and int(m.group(1)) > 1000 and int(m.group(2)) > int(m.group(1)), b[2][-400:]


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(os.path.join(REF, "lencod_b2t"))), reason="oracle/_ref/lencod_b2t not built")
@pytest.mark.parametrize("qp", [32])
def test_lencod_with_cuda_transform_quant_is_bit_identical(qp):
    """lencod_b2t: residual_transform_quant_luma_4x4 (block.c:660: forward4x4, quant_4x4_normal, inverse4x4, sample_reconstruct)
    of every luma 4x4 block the mode decision tries, plus the motion search, served by libb2me.so."""
    W, H, frames = 176, 144, 2
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=13))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=1, search_range=8, qp=qp)
        b = _encode("lencod_b2t", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=1, search_range=8, qp=qp,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 500
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) transform/quant calls", b[2])
        assert m and int(m.group(1)) > 10000, b[2][-400:]


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(os.path.join(REF, "lencod_b2q"))), reason="oracle/_ref/lencod_b2q not built")
@pytest.mark.parametrize("extra,what", [
    (("SymbolMode=1", "Transform8x8Mode=1"), "8x8"),               # CABAC: residual_transform_quant_luma_8x8's 64-coefficient list
    ((), "chroma"),                                                 # CAVLC, 4x4 transform: luma 4x4, Intra16x16 and chroma paths
])
def test_lencod_with_every_residual_path_on_the_gpu_is_bit_identical(extra, what):
    """lencod_b2q: residual_transform_quant_luma_4x4 / _8x8 / _16x16 and residual_transform_quant_chroma_4x4 (block.c:660, 207, 953;
    transform8x8.c:522) served by b2tq_4x4 / b2tq_8x8 / b2tq_16x16 / b2tq_chroma under the stock motion search and mode decision."""
    W, H, frames = 176, 144, 2
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=17))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=1, search_range=8, qp=30, search_mode=3, extra=extra)
        b = _encode("lencod_b2q", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=1, search_range=8, qp=30, search_mode=3, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 500
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) transform/quant calls, (\d+) 8x8, (\d+) Intra16x16, (\d+) chroma", b[2])
        assert m, b[2][-400:]
        n4, n8, n16, nc = (int(x) for x in m.groups())
        assert n16 > 100 and nc > 1000 and (n8 > 1000 if what == "8x8" else n4 > 10000), m.group(0)


@pytest.mark.skipif(not have, reason="oracle/_ref/lencod_b2 not built")
def test_dropin_fails_loudly_without_a_gpu():
    """No CPU fallback behind the boundary: without a CUDA device the shim stops the encoder."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    W, H = 176, 144
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, 2, seed=1))
        with pytest.raises(RuntimeError, match="b2me"):
            jm_run.run_lencod(yuv, W, H, 2, os.path.join(d, "b2"), exe="lencod_b2")


@pytest.mark.gpu
@pytest.mark.skipif(not (have and os.path.exists(os.path.join(REF, "lencod_b2f"))), reason="oracle/_ref/lencod_b2f not built")
@pytest.mark.parametrize("frames,nrefs,sr,qp,bframes", [(6, 2, 16, 28, 0), (5, 2, 8, 30, 1)])
def test_fast_full_search_with_gpu_sad_tables_is_bit_identical(frames, nrefs, sr, qp, bframes):
    """lencod_b2f, SearchMode 0: me_fullfast.o AND me_fullsearch.o out of the link; setup_fast_full_search's SAD tables come from
    b2me_sad_table (one GPU call per macroblock and reference), fast_full_search_motion_estimation scans them in the shim with the
    max_mvd guard (JM/lencod/src/me_fullfast.c:618-689)."""
    W, H = 176, 144
    extra = ("RestrictSearchRange=2",)
    if bframes:
        extra += ("NumberBFrames=1", "HierarchicalCoding=0", "BReferencePictures=0", "QPBSlice=30", "BList1References=1")
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "in.yuv")
        open(yuv, "wb").write(synth.yuv420_sequence(W, H, frames, seed=21))
        a = _encode("lencod", yuv, W, H, frames, os.path.join(d, "stock"), nrefs=nrefs, search_range=sr, qp=qp, search_mode=0, extra=extra)
        b = _encode("lencod_b2f", yuv, W, H, frames, os.path.join(d, "b2"), nrefs=nrefs, search_range=sr, qp=qp, search_mode=0, extra=extra,
                    env={"B2ME_SHIM_VERBOSE": "1"})
        assert len(a[0]) > 1000
        assert a[0] == b[0], "bitstreams differ"
        assert a[1] == b[1], "reconstructions differ"
        import re
        m = re.search(r"(\d+) fast-full-search set-ups \(GPU\), (\d+) table scans", b[2])
        assert m and int(m.group(1)) > 100 and int(m.group(2)) > 30 * int(m.group(1)), b[2][-400:]
