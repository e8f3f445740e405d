"""CPU: host-side mirror of the reference interface — module layout, checkpoint keys, KG table
index quirk, input contract, gradient-buffer layout, sharding."""
import numpy as np
import pytest
import torch

from _util import build_model, load_fixture, seeded_weights
from oracle import weights
from stonkgs_b200 import StkError, synthetic
from stonkgs_b200.embeddings import shard_bounds


@pytest.fixture(scope="module")
def small():
    fix, meta, batch = load_fixture("L2_B2_N997")
    sd, rows = seeded_weights(meta)
    return fix, meta, batch, sd, rows, build_model(meta, sd, rows)


def test_state_dict_layout_matches_reference(small):
    _, meta, _, sd, _, model = small
    assert sorted(model.state_dict().keys()) == sorted(sd.keys())      # exactly the reference key set
    assert len(weights.all_keys(175003, 12)) == 413                    # 413 keys at 12 layers
    assert model.config.kg_vocab_size == meta["n_kg"]
    assert model.cls.predictions.half_length == 256
    assert all(not p.requires_grad for p in model.lm_backbone.parameters())
    assert model.cls.predictions.text_decoder.bias is None and model.cls.predictions.entity_decoder.bias is None


def test_kg_table_index_quirk(small):
    fix, meta, _, _, rows, model = small
    n = meta["n_kg"]
    assert model.kg_table.shape == (n + 3, 768)
    # file row j sits at numeric_indices[j] = j-th element of range(N+3) \ {100,102,103}
    assert torch.equal(model.kg_table[99], torch.from_numpy(rows[99]))
    assert torch.equal(model.kg_table[101], torch.from_numpy(rows[100]))   # shift 1 at 101
    assert torch.equal(model.kg_table[104], torch.from_numpy(rows[101]))   # shift 3 from 104 on
    assert torch.equal(model.kg_table[n + 2], torch.from_numpy(rows[n - 1]))
    ids = [int(v) for v in fix["kg_probe_ids"]]
    normal = [i for i, v in enumerate(ids) if v not in (100, 102, 103)]
    assert np.array_equal(model.kg_table[ids].numpy()[normal], fix["kg_probe_rows"][normal])
    assert model.kg_idx_to_name[104] == "n101" and len(model.kg_backbone) == n + 3
    with pytest.raises(KeyError):
        model.kg_backbone[n + 3]


def test_out_of_table_id_raises_keyerror_like_reference(small):
    _, meta, batch, _, _, model = small
    bad = batch["input_ids"].clone()
    bad[0, 300] = meta["n_kg"] + 3
    with pytest.raises(KeyError):
        model._check_ids(bad)


def test_no_cpu_fallback(small):
    _, _, batch, _, _, model = small
    with pytest.raises(StkError):
        model(**batch)


def test_synthetic_batch_follows_input_contract():
    b = synthetic.make_batch(4, 5000, seed=3)
    ids, mask, tt = b["input_ids"], b["attention_mask"], b["token_type_ids"]
    assert ids.shape == (4, 512) and ids.dtype == torch.int64
    assert (ids[:, 0] == 101).all() and (ids[:, 256 + 127] == 102).all() and (ids[:, 511] == 102).all()
    assert (mask[:, 256:] == 1).all() and (tt[:, :256] == 0).all() and (tt[:, 256:] == 1).all()
    lens = mask[:, :256].sum(1)
    for r in range(4):
        assert ids[r, lens[r] - 1] == 102 and (ids[r, lens[r]:256] == 0).all()
    assert ((b["masked_lm_labels"] != -100).sum(1) == 38).all() and ((b["ent_masked_lm_labels"] != -100).sum(1) == 38).all()
    assert ids[:, 256:].max() < 5003


def test_grad_buffer_layout(small):
    from stonkgs_b200.training import GradBuffer
    _, _, _, _, _, model = small
    gb = GradBuffer(model)
    live = {id(p) for p, _ in gb.param_views}
    from oracle import stonkgs_oracle as orc
    named = dict(model.named_parameters())
    assert live == {id(named[k]) for k in orc.live_keys(dict(model.state_dict()))}
    assert gb.entries[0] == "w_ent" and gb.entries[-1] == "emb_b"          # reverse execution order
    for name, (off, n, shape) in gb.offsets.items():
        assert off % 4 == 0                                                # 16-byte aligned segments (TMA)
    q = model.bert.encoder.layer[0].attention.self
    vq = [v for p, v in gb.param_views if p is q.query.weight][0]
    vk = [v for p, v in gb.param_views if p is q.key.weight][0]
    assert vk.data_ptr() - vq.data_ptr() == 768 * 768 * 4                  # q|k|v grads are one fused block
    assert gb.prepare() == "fresh"
    gb.publish("fresh")
    assert q.query.weight.grad.data_ptr() == vq.data_ptr() and gb.prepare() == "accumulate"
    model.zero_grad(set_to_none=True)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_checkpoint_round_trip_hf_layout(tmp_path, small):
    """save_pretrained -> from_pretrained (offline) keeps every weight; the HF config carries kg_vocab_size
    (reference stonkgs_model.py:96-97) and extra kwargs reach __init__ like in the reference (api/api.py:107-110)."""
    import json
    from stonkgs_b200.model import STonKGsForPreTraining
    _, meta, _, sd, rows, model = small
    model.save_pretrained(tmp_path)
    cfg = json.load(open(tmp_path / "config.json"))
    assert cfg["kg_vocab_size"] == meta["n_kg"] and cfg["architectures"] == ["STonKGsForPreTraining"]
    again = STonKGsForPreTraining.from_pretrained(tmp_path, kg_embedding_dict_path=rows)
    sd2 = again.state_dict()
    assert sorted(sd2) == sorted(sd) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert again.kg_table.device.type == "cpu" and torch.equal(again.kg_table, model.kg_table)
    # the reference's legacy format: a plain state dict with all aliased keys
    torch.save(model.state_dict(), tmp_path / "pytorch_model.bin")
    fresh = build_model(meta, weights.make_state_dict(meta["n_kg"], meta["layers"], seed=5), rows)
    missing, unexpected = fresh.load_state_dict(torch.load(tmp_path / "pytorch_model.bin"), strict=True)
    assert not missing and not unexpected


def test_frozen_live_parameters_are_left_alone(small):
    """requires_grad=False on a live parameter (frozen lower layers while fine-tuning): it keeps its slot in the flat
    buffer but gets no .grad, is not handed to the optimizer and is not part of the clipping norm (torch semantics)."""
    _, meta, _, _, _, model = small
    model._grad_buffer = None
    full = model.grad_buffer()
    n_views = len(full.param_views)
    assert full.trainable_runs() == [(0, full.flat.numel())] and not full.stale()
    layer0 = list(model.bert.encoder.layer[0].parameters())
    try:
        for p in layer0:
            p.requires_grad_(False)
        assert full.stale()
        gb = model.grad_buffer()                                   # rebuilt: the layout is unchanged, the views are filtered
        assert gb is not full and gb.offsets == full.offsets
        assert len(gb.param_views) == n_views - len(layer0)
        frozen = {id(p) for p in layer0}
        assert not any(id(p) in frozen for p, _ in gb.param_views)
        runs = gb.trainable_runs()
        assert len(runs) == 2 and runs[0][0] == 0 and runs[-1][1] == gb.flat.numel()       # layer 0 sits before the embeddings
        assert runs[1][0] - runs[0][1] >= sum(p.numel() for p in layer0)                   # the frozen layer's segments
        gb.publish("fresh")
        assert all(p.grad is None for p in layer0) and model.bert.pooler.dense.weight.grad is not None
    finally:
        for p in layer0:
            p.requires_grad_(True)
        model.zero_grad(set_to_none=True)
        model._grad_buffer = None


def test_bucket_plan_with_a_short_tail():
    """dp.plan_buckets: contiguous cover in production order, a segment above the target ships alone, and the last
    bucket — the only one whose all-reduce cannot hide behind backward — is cut to the trailing segments that fit."""
    from stonkgs_b200.dp import plan_buckets
    entries, offsets, tot = [], {}, 0
    for name, k in [("w_ent", 1000), ("w_text", 300), ("a", 90), ("b", 3), ("c", 120), ("d", 50), ("e", 50), ("f", 7)]:
        entries.append(name)
        offsets[name] = (tot, k, (k,))
        tot += (k + 7) // 8 * 8
    for tail in (0, 70, 10 ** 6):
        bs = plan_buckets(entries, offsets, 200, tail)
        assert bs[0].names == ["w_ent"] and bs[0].start == 0 and bs[-1].end == tot
        assert all(x.end == y.start for x, y in zip(bs, bs[1:]))
        assert [n for b in bs for n in b.names] == entries
        assert all(b.start % 8 == 0 for b in bs)                    # bf16 wire slices stay 16-byte aligned
    assert plan_buckets(entries, offsets, 200, 70)[-1].names == ["e", "f"]           # 56 + 8 <= 70, + 56 would not fit
    assert plan_buckets(entries, offsets, 200, 1)[-1].names == ["f"]                 # at least one segment
    assert len(plan_buckets(entries, offsets, 200, 10 ** 6)) == 2                    # a tail never swallows the first segment


def test_label_capacity_rules(small):
    from stonkgs_b200.training import label_capacity
    _, _, batch, _, _, model = small
    mlm = batch["masked_lm_labels"]
    assert label_capacity(model, mlm, 256) == int((mlm != -100).sum())              # host labels: counted exactly

    class _OnDevice:                                                                 # stands in for a CUDA tensor
        is_cuda, shape = True, (8, 256)

    assert label_capacity(model, _OnDevice(), 256) == 38 * 8                         # int(0.15 * 256) per pair (reference)
    assert label_capacity(model, _OnDevice(), 4) == 4 * 8                            # TransE entity part: every position
    model.label_capacity = 50
    try:
        assert label_capacity(model, _OnDevice(), 256) == 50 * 8 and label_capacity(model, _OnDevice(), 4) == 4 * 8
    finally:
        model.label_capacity = None


def test_lazy_prediction_logits_behave_like_the_pair(monkeypatch):
    """prediction_logits of a training step (training.LazyPredictionLogits): nothing is computed until the pair is
    looked at; indexing, unpacking, len() and .detach() all see the reference's (text, entity) pair, and HF Trainer's
    nested_detach passes the object through (it only rebuilds lists / tuples / mappings / tensors)."""
    from transformers.trainer_pt_utils import nested_detach
    from stonkgs_b200 import training
    calls = []

    def fake_dense(hw, seq, B, shape):
        calls.append(B)
        return torch.ones(B, 256, 7), torch.zeros(B, 256, 5)

    monkeypatch.setattr(training, "dense_prediction_logits", fake_dense)
    lazy = training.LazyPredictionLogits("hw", "seq", 3, "shape")
    assert not calls and len(lazy) == 2 and "pending" in repr(lazy)
    text, ent = lazy                                           # unpacking materialises, once
    assert calls == [3] and text.shape == (3, 256, 7) and ent.shape == (3, 256, 5)
    assert lazy[0] is text and lazy[1] is ent and calls == [3]
    out = nested_detach((torch.zeros(()), lazy, torch.zeros(3, 2)))
    assert torch.equal(out[1][0], text) and torch.equal(out[1][1], ent) and calls == [3]
    det = lazy.detach()
    assert isinstance(det, tuple) and torch.equal(det[0], text)


def test_get_stonkgs_embeddings_empty_frame():
    """An empty frame gives the reference's empty result (``pd.DataFrame(columns=["embedding"])``, stonkgs_for_embeddings.py:164)."""
    import pandas as pd
    from stonkgs_b200.embeddings import get_stonkgs_embeddings
    df = pd.DataFrame({"input_ids": [], "attention_mask": [], "token_type_ids": []})
    seen = []

    def fake_embed(i, m, t):
        seen.append(i.shape[0])
        return np.zeros((0, 768), dtype=np.float32)

    out = get_stonkgs_embeddings(df, _embed_fn=fake_embed)
    assert list(out.columns) == ["embedding"] and len(out) == 0 and seen == [0]


def test_plan_live_rows():
    """engine.plan_live_rows (skip_padding extraction): every pair's attended rows come first ([CLS] in front, original
    order kept), the pair is cut at the first multiple of 128 rows that holds them, pairs are grouped by that length; a pair
    without any attended key keeps all rows."""
    from stonkgs_b200 import engine, synthetic
    batch = synthetic.make_batch(37, 997, seed=2, with_labels=False)
    mask = batch["attention_mask"].numpy().copy()
    mask[3] = 0                     # nothing attended: the reference's uniform attention needs every key
    mask[4, :] = 1                  # no padding at all
    mask[5, 0] = 0                  # [CLS] masked as a key: still the first packed row (it is read as a query)
    mask[6, 10:256] = 0
    mask[6, 300:400] = 0            # holes in the KG half: 10 + 156 = 166 attended rows -> 256
    mask[7, 1:256] = 0              # 1 + 256 attended rows: one more than two tiles -> 384
    S = mask.shape[1]
    plan = engine.plan_live_rows(mask)
    lengths = [sb for sb, _, _ in plan]
    assert lengths == sorted(set(lengths), reverse=True) and set(lengths) <= {128, 256, 384, 512}
    seen = {}
    for sb, pairs, rows in plan:
        assert pairs.dtype == np.int64 and rows.dtype == np.int32 and rows.shape == (len(pairs) * sb,)
        for b, r in zip(pairs.tolist(), rows.reshape(len(pairs), sb)):
            pos = r - b * S
            assert pos[0] == 0 and len(set(pos.tolist())) == sb and pos.min() >= 0 and pos.max() < S
            attended = set(np.nonzero(mask[b])[0].tolist())
            if sb == S:
                assert list(pos) == list(range(S))                             # nothing cut: the pair is left as it is
            if attended:
                n_live = len(attended | {0})
                assert sb == min(S, -(-n_live // 128) * 128)
                assert attended | {0} <= set(pos.tolist())                   # every attended row is kept ...
                if sb < S:
                    assert set(pos[:n_live].tolist()) == attended | {0}      # ... in front ...
                    assert list(pos[1:n_live]) == sorted(pos[1:n_live])        # ... in its original order
            else:
                assert sb == S and list(pos) == list(range(S))
            seen[b] = sb
    assert sorted(seen) == list(range(37))
    assert seen[3] == 512 and seen[4] == 512 and seen[6] == 256 and seen[7] == 384
    # a sequence length that is not a multiple of the tile keeps every pair whole
    assert [sb for sb, _, _ in engine.plan_live_rows(np.ones((3, 260), dtype=np.int64))] == [260]
