"""CPU checks that pin the oracle and the plan compiler (no GPU).

* The reference ships no golden vectors for this path, so the anchors are (a) the decodes of the
  reference's own sample video frames recorded in SURVEY.md Appendix E and (b) tests/golden/golden.json, produced by
  the graph-level oracle (oracle/graph_interp.py runs the shipped inference.pdmodel op by op).
* The packed plans (what the CUDA engine executes) must reproduce those results on the CPU plan interpreter: that
  proves fusion / BN folding / concat aliasing / weight re-layout preserve the arithmetic of the shipped graphs.
"""
import json
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import hostlogic as hl
from oracle.pipeline import OraclePipeline
from video_subtitle_extractor_b200 import plan as P
from video_subtitle_extractor_b200 import weights

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "backend", "models")), reason="reference tree not present")


@pytest.fixture(scope="module")
def plan_oracle():
    return OraclePipeline.from_plans(weights.load_plan_blob("V4/ch_det_fast"), weights.load_plan_blob("V4/en_rec_fast"))


def test_plan_oracle_reproduces_golden(plan_oracle):
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        golden = json.load(f)
    assert len(golden["cases"]) >= 4
    for c in golden["cases"]:
        frame = cv2.imread(os.path.join(GOLDEN, c["image"]))
        assert list(frame.shape) == c["shape"]
        r = plan_oracle.ocr(frame)
        assert [np.asarray(b).astype(int).tolist() for b in r.boxes] == c["boxes"]
        assert r.ids == c["ids"]
        assert r.rec_widths == c["rec_widths"]
        assert np.allclose(r.scores, c["rec_scores"], atol=1e-4)
        assert np.allclose(r.det_scores, c["det_scores"], atol=1e-4)
        assert plan_oracle.detect(frame).astype(int).tolist() == c["det_only_boxes"]


def test_golden_text_is_what_the_video_shows():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        golden = json.load(f)
    texts = {c["image"]: [hl.ids_to_text(i, hl.EN_CHARACTERS) for i in c["ids"]] for c in golden["cases"]}
    assert texts["test_en_300.png"] == ["As far as we can go."]      # SURVEY.md Appendix E anchor
    assert texts["test_en_2500.png"] == []


def test_plan_blob_round_trip():
    for name in weights.DEFAULT_MODELS:
        blob = weights.load_plan_blob(name)
        assert P.deserialize(blob).serialize() == blob


@needs_ref
def test_graph_oracle_matches_survey_anchor():
    orc = OraclePipeline(f"{REF}/backend/models/V4/ch_det_fast", f"{REF}/backend/models/V4/en_rec_fast")
    cap = cv2.VideoCapture(f"{REF}/test/test_en.mp4")
    cap.set(cv2.CAP_PROP_POS_FRAMES, 300)
    ok, frame = cap.read()
    assert ok
    r = orc.ocr(frame)
    # SURVEY.md Appendix E recorded these with a throw-away interpreter whose unclip was "grow the rectangle by d"
    # (SURVEY D.3), so boxes may differ by a pixel and the crop-dependent text by one symbol.
    texts = [hl.ids_to_text(i, hl.EN_CHARACTERS) for i in r.ids]
    assert texts[0] == "Yami Sukehiro"
    assert texts[1].replace(" ", "") == "Asfaraswecango."
    boxes = [np.asarray(b).astype(int) for b in r.boxes]
    assert np.abs(boxes[1] - np.array([[454, 642], [820, 649], [819, 689], [453, 682]])).max() <= 1
    assert np.abs(boxes[0] - np.array([[979, 31], [1222, 31], [1222, 60], [979, 60]])).max() <= 1
    assert r.scores[0] > 0.95 and r.scores[1] > 0.92


@needs_ref
@pytest.mark.parametrize("name,shape", [("V4/ch_det_fast", (1, 3, 96, 160)), ("V4/en_rec_fast", (2, 3, 48, 336)),
                                        ("V4/ch_rec_fast", (1, 3, 48, 320)), ("V3/japan_rec_fast", (1, 3, 48, 320)),
                                        ("V3/korean_rec_fast", (1, 3, 48, 328)), ("V2/ch_rec", (2, 3, 32, 168)),
                                        # the accurate-mode (server) models: PP-HGNet + LK-PAN + PFHeadLocal (whose concat slices
                                        # the plan compiler reorders, plan.py::_reorder_concats) and the 6625-class SVTR recogniser
                                        ("V4/ch_det", (1, 3, 96, 128)), ("V4/ch_rec", (1, 3, 48, 160))])
def test_compiled_plan_equals_shipped_graph(name, shape):
    from oracle.graph_interp import GraphInterpreter
    from oracle.plan_interp import PlanInterpreter
    from video_subtitle_extractor_b200.loader import load_model
    model = load_model(f"{REF}/backend/models/{name}")
    plan = P.compile_model(model, name=name, fetch_cols=[0])
    x = torch.from_numpy(np.random.default_rng(0).standard_normal(shape).astype(np.float32))
    ref = GraphInterpreter(model).run(x)[0]
    got = PlanInterpreter(P.deserialize(plan.serialize())).run(x)[0]
    if ref.dim() == 3:   # [B, T, C] vs plan layout [B, C, 1, T]
        got = got[:, :, 0, :].permute(0, 2, 1)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))
    assert len(plan.steps) < 0.45 * len(model.program.ops)   # fusion actually happened


@needs_ref
def test_all_shipped_models_compile():
    from video_subtitle_extractor_b200.loader import load_model
    root = f"{REF}/backend/models"
    n = 0
    for ver in sorted(os.listdir(root)):
        for name in sorted(os.listdir(os.path.join(root, ver))):
            d = os.path.join(root, ver, name)
            if not os.path.exists(os.path.join(d, "inference.pdmodel")):
                continue
            plan = P.compile_model(load_model(d), name=f"{ver}/{name}", fetch_cols=[0])
            assert plan.steps and plan.output_vids
            n += 1
    assert n == 21


@needs_ref
def test_rnn_oracle_matches_torch_lstm():
    """Paddle's `rnn` op (V2/ch_rec op#140: 2-layer bidirectional LSTM, gate order i,f,g,o, WeightList = all weights then all
    biases) as restated in oracle/graph_interp.py::_rnn, against torch.nn.LSTM loaded with the same shipped weights — an
    independent implementation of the same recurrence."""
    from oracle.graph_interp import GraphInterpreter
    from video_subtitle_extractor_b200.loader import load_model
    model = load_model(f"{REF}/backend/models/V2/ch_rec")
    gi = GraphInterpreter(model)
    op = [o for o in model.program.ops if o.type == "rnn"][0]
    a = op.attrs
    assert (a["mode"], a["hidden_size"], a["num_layers"], a["is_bidirec"]) == ("LSTM", 256, 2, True)
    wl = [gi._get({}, n) for n in op.inputs["WeightList"]]
    lstm = torch.nn.LSTM(a["input_size"], a["hidden_size"], a["num_layers"], bidirectional=True)
    nw = 4
    with torch.no_grad():
        for layer in range(2):
            for d, suffix in enumerate(("", "_reverse")):
                k = layer * 2 + d
                getattr(lstm, f"weight_ih_l{layer}{suffix}").copy_(wl[2 * k])
                getattr(lstm, f"weight_hh_l{layer}{suffix}").copy_(wl[2 * k + 1])
                getattr(lstm, f"bias_ih_l{layer}{suffix}").copy_(wl[2 * nw + 2 * k])
                getattr(lstm, f"bias_hh_l{layer}{suffix}").copy_(wl[2 * nw + 2 * k + 1])
        x = torch.from_numpy(np.random.default_rng(3).standard_normal((37, 3, a["input_size"])).astype(np.float32))
        want, _ = lstm(x)
        got = gi._rnn(op, {op.inp("Input"): x})
    assert got.shape == want.shape == (37, 3, 512)
    assert float((got - want).abs().max()) < 1e-5
