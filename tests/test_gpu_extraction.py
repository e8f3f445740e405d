"""GPU: embedding extraction at the BASELINE shape (12 layers, N_kg = 175 003, batch 256) through
size-independent properties, plus the DataFrame API of get_stonkgs_embeddings."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_model():
    from transformers import BertConfig
    from oracle import weights
    from stonkgs_b200.model import STonKGsForPreTraining
    n_kg = 175003
    torch.manual_seed(0)
    model = STonKGsForPreTraining(None, BertConfig(vocab_size=28996), weights.make_kg_table(n_kg, 0))
    return model.eval().to("cuda"), n_kg


def test_batch_256_properties(full_model):
    from stonkgs_b200 import synthetic
    model, n_kg = full_model
    batch = synthetic.make_batch(256, n_kg, seed=11, with_labels=False)
    e = model.embed(**batch)
    assert e.shape == (256, 768) and e.dtype == torch.float32 and torch.isfinite(e).all()
    assert e.abs().max() <= 1.0                                   # tanh pooler (HF:456-468)
    # a pair's embedding does not depend on what else is in the batch: bit-exact on sub-batches
    sub = {k: v[40:72] for k, v in batch.items()}
    assert torch.equal(model.embed(**sub), e[40:72])
    one = {k: v[255:256] for k, v in batch.items()}
    assert torch.equal(model.embed(**one), e[255:256])
    # permutation equivariance and idempotence
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(0))
    assert torch.equal(model.embed(**{k: v[perm] for k, v in batch.items()}), e[perm.cuda()])
    assert torch.equal(model.embed(**batch), e)
    # editing one pair leaves every other pair's embedding untouched, bit for bit
    lens = batch["attention_mask"][:, :256].sum(1)
    b = int((lens < 250).nonzero()[0])
    edited = {k: v.clone() for k, v in batch.items()}
    edited["input_ids"][b, 255] = 1234
    e2 = model.embed(**edited)
    others = [i for i in range(256) if i != b]
    assert torch.equal(e2[others], e[others])


def test_get_stonkgs_embeddings_dataframe_api(full_model):
    from stonkgs_b200 import get_stonkgs_embeddings, synthetic
    model, n_kg = full_model
    batch = synthetic.make_batch(37, n_kg, seed=5)   # label columns present, like the reference's rows
    df = pd.DataFrame({k: list(v.numpy()) for k, v in batch.items()}, index=[f"r{i}" for i in range(37)])
    out = get_stonkgs_embeddings(df, model=model, batch_size=16)
    assert list(out.columns) == ["embedding"] and list(out.index) == list(df.index)
    arr = np.asarray(out["embedding"].tolist(), dtype=np.float32)
    ref = model.embed(batch["input_ids"], batch["attention_mask"], batch["token_type_ids"]).cpu().numpy()
    assert np.array_equal(arr, ref)
    some = get_stonkgs_embeddings(df, list_of_indices=["r3", "r20"], model=model)
    assert np.array_equal(np.asarray(some["embedding"].tolist(), dtype=np.float32), ref[[3, 20]])
