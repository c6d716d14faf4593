"""Exemplar bank: the visual exemplars the teacher collects per caption noun (second half).

The reference keeps a Python dict `word -> {'emb', 'quality', 'score', 'info'}` ('SINGLE') or
`word -> {'emb', 'accum_quality', 'accum_score'}` ('ACCUM') on every rank
(modeling/detector/st_generalized_rcnn.py:107-132), pickles it per rank
(`exemplars_{rank}_{type}.pkl`, :134-137), and at start-up unpickles and merges every rank's file
(:139-162); `combine_embs` then adds `lambda_exemplar * exemplar` to the word embeddings of the nouns that
have one (:164-178).  Here:

  * the bank is three flat arrays (word index into a fixed vocabulary, embedding matrix [V, D], quality /
    score vectors) that live on the device, so `update` is a batched scatter (one pass for a whole image
    batch instead of a Python loop over nouns with `.cpu()` per field) and `combine_embs` is one fused
    gather-add-normalise;
  * the file is a flat, versioned, little-endian container (same family as records.py: 64-byte header, JSON
    meta with the vocabulary, then the arrays; np.memmap-able, no pickle);
  * `merge` implements both reference policies: SINGLE keeps the higher quality per word (:150-152), ACCUM is
    the quality-weighted mean with accumulated quality and score (the commented-out `combine_exemplar`, :91-105).

Host / torch code only; no CUDA kernel of its own.
"""
import json
import os
import struct

import numpy as np
import torch

MAGIC = b"B2EX"
VERSION = 1
_HEADER = struct.Struct("<4sIIIIII36x")   # magic, version, n_words, dim, type (0 SINGLE / 1 ACCUM), n_valid, meta_len
assert _HEADER.size == 64
TYPES = ("SINGLE", "ACCUM")
_EPS = 1e-5   # combine_exemplar's eps (st_generalized_rcnn.py:94)


def _pad64(n):
    return (n + 63) // 64 * 64


class ExemplarBank(object):
    def __init__(self, vocab, dim, exemplar_type="SINGLE", device="cpu"):
        if exemplar_type not in TYPES:
            raise ValueError("Unknown exemplar type")     # the reference's own message (:131)
        self.vocab = list(vocab)
        self.index = {w: i for i, w in enumerate(self.vocab)}
        if len(self.index) != len(self.vocab):
            raise ValueError("vocabulary has duplicate words")
        self.type = exemplar_type
        v = len(self.vocab)
        self.emb = torch.zeros((v, dim), dtype=torch.float32, device=device)
        self.quality = torch.zeros((v,), dtype=torch.float32, device=device)   # SINGLE: quality; ACCUM: accum_quality
        self.score = torch.zeros((v,), dtype=torch.float32, device=device)     # SINGLE: score;   ACCUM: accum_score
        self.valid = torch.zeros((v,), dtype=torch.bool, device=device)

    # ------------------------------------------------------------------------------------------
    def __len__(self):
        return int(self.valid.sum())

    def __contains__(self, word):
        i = self.index.get(word)
        return i is not None and bool(self.valid[i])

    def get(self, word):
        """The reference's dict entry for `word` (or None)."""
        i = self.index.get(word)
        if i is None or not bool(self.valid[i]):
            return None
        if self.type == "SINGLE":
            return {"emb": self.emb[i].cpu(), "quality": self.quality[i].cpu(), "score": self.score[i].cpu()}
        return {"emb": self.emb[i].cpu(), "accum_quality": self.quality[i].cpu(), "accum_score": self.score[i].cpu()}

    def word_ids(self, words):
        """Vocabulary indices of `words`; unknown words raise (the vocabulary is fixed per run: the LVIS noun
        list of st_generalized_rcnn.py:71-75)."""
        try:
            return torch.tensor([self.index[w] for w in words], dtype=torch.int64, device=self.emb.device)
        except KeyError as e:
            raise KeyError("word %s is not in the exemplar vocabulary" % e)

    # ------------------------------------------------------------------------------------------
    def update(self, word_ids, embs, scores, consistencies=None):
        """Batched `update_exemplars` (:107-132) for any number of pseudo-labels at once (one image or a
        whole batch).  word_ids int64 [N]; embs [N, D]; scores, consistencies [N].  Within one call several
        labels may name the same word: SINGLE keeps the best of them, ACCUM folds all of them in, exactly
        as the reference's sequential loop does."""
        dev = self.emb.device
        word_ids = torch.as_tensor(word_ids, dtype=torch.int64, device=dev)
        if word_ids.numel() == 0:
            return
        embs = torch.nn.functional.normalize(torch.as_tensor(embs, dtype=torch.float32, device=dev), dim=-1)   # :117
        scores = torch.as_tensor(scores, dtype=torch.float32, device=dev)
        q = scores if consistencies is None else scores * torch.as_tensor(consistencies, dtype=torch.float32, device=dev)
        if self.type == "SINGLE":
            # a word's new entry = its highest-quality candidate, first one on ties (the loop replaces only on `<`)
            v = len(self.vocab)
            best_q = torch.full((v,), float("-inf"), device=dev).scatter_reduce(0, word_ids, q, "amax", include_self=True)
            is_best = q == best_q[word_ids]
            pos = torch.arange(word_ids.numel(), device=dev)
            first = torch.full((v,), word_ids.numel(), dtype=torch.int64, device=dev).scatter_reduce(
                0, word_ids[is_best], pos[is_best], "amin", include_self=True)
            words = torch.unique(word_ids)
            src = first[words]
            take = (~self.valid[words]) | (self.quality[words] < q[src])          # :121
            w, s = words[take], src[take]
            self.emb[w] = embs[s]
            self.quality[w] = q[s]
            self.score[w] = scores[s]
            self.valid[w] = True
        else:
            # sequential fold p = combine(p, new) over the candidates of a word == one weighted sum:
            # emb <- (emb * Q + sum_i emb_i q_i) / (Q + sum_i q_i + eps) applied label by label; done in order
            # here so that eps enters exactly as often as in the reference
            order = torch.argsort(word_ids, stable=True)
            for i in order.tolist():
                w = int(word_ids[i])
                if not bool(self.valid[w]):
                    self.emb[w], self.quality[w], self.score[w] = embs[i], q[i], scores[i]
                    self.valid[w] = True
                else:
                    aq = self.quality[w] + q[i]
                    self.emb[w] = (self.emb[w] * self.quality[w] + embs[i] * q[i]) / (aq + _EPS)
                    self.quality[w] = aq
                    self.score[w] = self.score[w] + scores[i]

    def merge(self, other):
        """Fold another bank of the same vocabulary in: what load_exemplars does with each rank's file (:147-158)."""
        if other.vocab != self.vocab or other.type != self.type or other.emb.shape != self.emb.shape:
            raise ValueError("exemplar banks differ in vocabulary, type or dimension")
        o_valid = other.valid.to(self.emb.device)
        o_emb, o_q, o_s = (t.to(self.emb.device) for t in (other.emb, other.quality, other.score))
        if self.type == "SINGLE":
            take = o_valid & ((~self.valid) | (self.quality < o_q))
            self.emb[take], self.quality[take], self.score[take] = o_emb[take], o_q[take], o_s[take]
        else:
            new = o_valid & ~self.valid
            both = o_valid & self.valid
            aq = self.quality + o_q
            comb = (self.emb * self.quality[:, None] + o_emb * o_q[:, None]) / (aq[:, None] + _EPS)
            self.emb[both] = comb[both]
            self.quality[both] = aq[both]
            self.score[both] = (self.score + o_s)[both]
            self.emb[new], self.quality[new], self.score[new] = o_emb[new], o_q[new], o_s[new]
        self.valid |= o_valid
        return self

    def combine_embs(self, word_ids, embs, lambda_exemplar):
        """`combine_embs` (:164-178): normalise(embs + lambda * exemplar) for the words that have an exemplar,
        normalise(embs) for the others; with an empty bank just normalise(embs).  Differentiable in
        `lambda_exemplar` (the reference keeps it in the graph for every row, :175-176)."""
        embs = torch.as_tensor(embs)
        if len(self) == 0:
            return torch.nn.functional.normalize(embs, dim=-1)
        word_ids = torch.as_tensor(word_ids, dtype=torch.int64, device=self.emb.device)
        ex = (self.emb[word_ids] * self.valid[word_ids][:, None]).to(embs.device, embs.dtype)
        return torch.nn.functional.normalize(embs.detach().clone() + lambda_exemplar * ex, dim=-1)

    # ------------------------------------------------------------------------------------------
    def save(self, path, meta=None):
        """Flat container: header, JSON meta (vocabulary, type), valid mask (uint8), quality, score, embeddings
        -- all V rows, so a reader can np.memmap the embedding block.  Written to path + '.tmp' and renamed."""
        m = dict(meta or {})
        m["vocab"] = self.vocab
        m["type"] = self.type
        mb = json.dumps(m, sort_keys=True).encode("utf-8")
        v, d = self.emb.shape
        valid = self.valid.cpu().numpy().astype(np.uint8)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(_HEADER.pack(MAGIC, VERSION, v, d, TYPES.index(self.type), int(valid.sum()), len(mb)))
            f.write(mb.ljust(_pad64(len(mb)), b"\0"))
            f.write(valid.tobytes().ljust(_pad64(v), b"\0"))
            f.write(np.ascontiguousarray(self.quality.cpu().numpy(), "<f4").tobytes().ljust(_pad64(4 * v), b"\0"))
            f.write(np.ascontiguousarray(self.score.cpu().numpy(), "<f4").tobytes().ljust(_pad64(4 * v), b"\0"))
            f.write(np.ascontiguousarray(self.emb.cpu().numpy(), "<f4").tobytes())
        os.replace(tmp, path)

    @classmethod
    def load(cls, path, device="cpu"):
        size = os.path.getsize(path)
        with open(path, "rb") as f:
            raw = f.read(_HEADER.size)
            if len(raw) != _HEADER.size:
                raise ValueError("truncated exemplar file")
            magic, version, v, d, t, n_valid, meta_len = _HEADER.unpack(raw)
            if magic != MAGIC:
                raise ValueError("not an exemplar file (bad magic %r)" % magic)
            if version != VERSION:
                raise ValueError("exemplar file version %d, this reader understands %d" % (version, VERSION))
            if t >= len(TYPES):
                raise ValueError("Unknown exemplar type")
            expect = _HEADER.size + _pad64(meta_len) + _pad64(v) + 2 * _pad64(4 * v) + 4 * v * d
            if size != expect:
                raise ValueError("exemplar file size does not match its header (truncated or corrupt)")
            meta = json.loads(f.read(_pad64(meta_len))[:meta_len].decode("utf-8"))
            valid = np.frombuffer(f.read(_pad64(v))[:v], dtype=np.uint8).astype(bool)
            quality = np.frombuffer(f.read(_pad64(4 * v))[: 4 * v], dtype="<f4").copy()
            score = np.frombuffer(f.read(_pad64(4 * v))[: 4 * v], dtype="<f4").copy()
            emb = np.frombuffer(f.read(4 * v * d), dtype="<f4").reshape(v, d).copy()
        if int(valid.sum()) != n_valid or len(meta.get("vocab", [])) != v:
            raise ValueError("exemplar file: header and payload disagree")
        bank = cls(meta["vocab"], d, TYPES[t], device)
        bank.emb, bank.quality, bank.score = (torch.from_numpy(a).to(device) for a in (emb, quality, score))
        bank.valid = torch.from_numpy(valid).to(device)
        return bank

    @classmethod
    def load_shards(cls, paths, device="cpu"):
        """load_exemplars (:139-162): merge every rank's file that exists; missing files are skipped like the
        reference's bare `except` does, anything else propagates."""
        bank = None
        for p in paths:
            if not os.path.exists(p):
                continue
            b = cls.load(p, device)
            bank = b if bank is None else bank.merge(b)
        return bank

    def to_reference_dict(self):
        """The dict the reference pickles (for interchange with an existing `exemplars_*.pkl` consumer)."""
        return {w: self.get(w) for w in self.vocab if w in self}

    @classmethod
    def from_reference_dict(cls, d, vocab, dim, exemplar_type="SINGLE", device="cpu"):
        bank = cls(vocab, dim, exemplar_type, device)
        qk, sk = ("quality", "score") if exemplar_type == "SINGLE" else ("accum_quality", "accum_score")
        for w, e in d.items():
            i = bank.index[w]
            bank.emb[i] = torch.as_tensor(e["emb"], dtype=torch.float32)
            bank.quality[i] = float(e[qk])
            bank.score[i] = float(e[sk])
            bank.valid[i] = True
        return bank
