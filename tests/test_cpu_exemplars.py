"""Exemplar bank (second half): the batched update / merge / combine_embs against a literal
restatement of the reference's dict code (st_generalized_rcnn.py:107-178), and the file round trip."""
import struct

import pytest
import torch
import torch.nn.functional as F

from cvpr22_cross_modal_pseudo_labeling_b200.exemplars import ExemplarBank

VOCAB = ["cat", "dog", "zebra", "traffic light", "surfboard", "kite"]
D = 16


def _ref_update(ex, typ, nns, scores, cons, embs):
    """update_exemplars, reference lines 107-132, on a plain dict."""
    q = scores * cons
    embs = F.normalize(embs, dim=-1)
    for i, nn in enumerate(nns):
        if typ == "SINGLE":
            pkg = {"emb": embs[i], "quality": q[i], "score": scores[i]}
            if nn not in ex or ex[nn]["quality"] < pkg["quality"]:
                ex[nn] = pkg
        else:
            pkg = {"emb": embs[i], "accum_quality": q[i], "accum_score": scores[i]}
            ex[nn] = pkg if nn not in ex else _ref_combine(ex[nn], pkg)


def _ref_combine(p1, p2):
    """combine_exemplar, reference lines 91-105."""
    aq = p1["accum_quality"] + p2["accum_quality"]
    return {"accum_quality": aq, "accum_score": p1["accum_score"] + p2["accum_score"],
            "emb": (p1["emb"] * p1["accum_quality"] + p2["emb"] * p2["accum_quality"]) / (aq + 1e-5)}


def _ref_combine_embs(ex, nns, embs, lam):
    """combine_embs, reference lines 164-178."""
    if len(ex) == 0:
        return F.normalize(embs, dim=-1)
    w = embs.clone().detach()
    for i, nn in enumerate(nns):
        w[i] = w[i] + (lam * ex[nn]["emb"] if nn in ex else lam * 0.0)
    return F.normalize(w, dim=-1)


def _batches(seed, n_batches=6):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        n = int(torch.randint(1, 9, (1,), generator=g))
        ids = torch.randint(0, len(VOCAB) - 1, (n,), generator=g)          # "kite" never appears
        out.append(([VOCAB[i] for i in ids.tolist()], torch.rand((n,), generator=g), torch.rand((n,), generator=g),
                    torch.randn((n, D), generator=g)))
    return out


@pytest.mark.parametrize("typ", ["SINGLE", "ACCUM"])
def test_update_matches_the_reference_loop(typ):
    bank, ref = ExemplarBank(VOCAB, D, typ), {}
    for nns, scores, cons, embs in _batches(3):
        bank.update(bank.word_ids(nns), embs, scores, cons)
        _ref_update(ref, typ, nns, scores, cons, embs)
    assert len(bank) == len(ref) and "kite" not in bank
    qk, sk = ("quality", "score") if typ == "SINGLE" else ("accum_quality", "accum_score")
    for w, e in ref.items():
        got = bank.get(w)
        assert torch.allclose(got["emb"], e["emb"], atol=1e-6) and torch.allclose(got[qk], e[qk]) and torch.allclose(got[sk], e[sk])
    # duplicates inside one call with equal quality: the first one wins (the loop replaces only on `<`)
    if typ == "SINGLE":
        b2 = ExemplarBank(VOCAB, D, typ)
        e = torch.randn((3, D))
        b2.update(b2.word_ids(["cat", "cat", "cat"]), e, torch.tensor([0.5, 0.9, 0.9]), torch.ones(3))
        assert torch.allclose(b2.get("cat")["emb"], F.normalize(e[1], dim=-1))


@pytest.mark.parametrize("typ", ["SINGLE", "ACCUM"])
def test_merge_of_rank_files_and_file_round_trip(tmp_path, typ):
    banks, ref = [], {}
    for r in range(3):
        b, d = ExemplarBank(VOCAB, D, typ), {}
        for nns, scores, cons, embs in _batches(10 + r, 3):
            b.update(b.word_ids(nns), embs, scores, cons)
            _ref_update(d, typ, nns, scores, cons, embs)
        p = str(tmp_path / ("exemplars_%d_%s.b2ex" % (r, typ)))
        b.save(p, meta={"rank": r})
        banks.append(p)
        for k in d:                                   # load_exemplars, reference lines 147-158
            if typ == "SINGLE":
                if k not in ref or ref[k]["quality"] < d[k]["quality"]:
                    ref[k] = d[k]
            else:
                ref[k] = d[k] if k not in ref else _ref_combine(ref[k], d[k])
    merged = ExemplarBank.load_shards(banks + [str(tmp_path / "missing.b2ex")])
    assert merged.type == typ and merged.vocab == VOCAB and len(merged) == len(ref)
    for w, e in ref.items():
        assert torch.allclose(merged.get(w)["emb"], e["emb"], atol=1e-6)
    one = ExemplarBank.load(banks[0])
    again = str(tmp_path / "again.b2ex")
    one.save(again)
    two = ExemplarBank.load(again)
    assert torch.equal(one.emb, two.emb) and torch.equal(one.valid, two.valid) and torch.equal(one.quality, two.quality)
    # interchange with the reference's pickled dict
    back = ExemplarBank.from_reference_dict(one.to_reference_dict(), VOCAB, D, typ)
    assert torch.equal(back.emb, one.emb) and torch.equal(back.valid, one.valid)


def test_combine_embs_and_gradient_of_lambda():
    bank, ref = ExemplarBank(VOCAB, D, "SINGLE"), {}
    embs0 = torch.randn((4, D))
    assert torch.allclose(bank.combine_embs(bank.word_ids(VOCAB[:4]), embs0, torch.tensor(0.3)), F.normalize(embs0, dim=-1))
    for nns, scores, cons, embs in _batches(5, 2):
        bank.update(bank.word_ids(nns), embs, scores, cons)
        _ref_update(ref, "SINGLE", nns, scores, cons, embs)
    nns = VOCAB                                        # includes "kite", which has no exemplar
    embs = torch.randn((len(nns), D))
    lam = torch.tensor(0.25, requires_grad=True)
    got = bank.combine_embs(bank.word_ids(nns), embs, lam)
    want = _ref_combine_embs(ref, nns, embs, lam.detach())
    assert torch.allclose(got, want, atol=1e-6)
    got.sum().backward()
    assert lam.grad is not None and float(lam.grad.abs()) > 0


def test_foreign_and_truncated_files_raise(tmp_path):
    b = ExemplarBank(VOCAB, D)
    b.update(b.word_ids(["cat"]), torch.randn((1, D)), torch.tensor([0.7]))
    p = str(tmp_path / "e.b2ex")
    b.save(p)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-8])
    with pytest.raises(ValueError):
        ExemplarBank.load(p)
    open(p, "wb").write(b"NOPE" + raw[4:])
    with pytest.raises(ValueError):
        ExemplarBank.load(p)
    open(p, "wb").write(raw[:4] + struct.pack("<I", 99) + raw[8:])
    with pytest.raises(ValueError):
        ExemplarBank.load(p)
    with pytest.raises(ValueError):
        ExemplarBank(VOCAB, D, "OTHER")
    with pytest.raises(KeyError):
        b.word_ids(["unicorn"])
