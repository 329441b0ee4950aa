"""Seeded synthetic inputs: the stand-ins for everything the hot path consumes that is not available offline
(datasets, ground truth, guide images, the bert-base-uncased vocabulary).  Used by the synthetic driver, by bench.py,
by tests/golden/make_golden.py (which runs the REFERENCE's own functions on them in the build container) and by the
tests (which run the oracle and the CUDA path on the same inputs).  Nothing here reads /root/reference."""
import zlib

import numpy as np
import torch

SEP_ID, CLS_ID, PAD_ID, ENC_ID = 102, 101, 0, 30523


class SyntheticWordPieceTokenizer:
    """Deterministic stand-in for BertTokenizer('bert-base-uncased') -- the vocab file is not available
    offline.  Words longer than `split_len` are broken into a head piece and
    '##' continuation pieces so the token->class merge (DRV:810-853) is exercised.  SEP=102, CLS=101, PAD=0."""

    def __init__(self, split_len=7, piece_len=4):
        self.split_len, self.piece_len = split_len, piece_len
        self._id2piece = {CLS_ID: "[CLS]", SEP_ID: "[SEP]", PAD_ID: "[PAD]", ENC_ID: "[ENC]"}
        self._piece2id = {v: k for k, v in self._id2piece.items()}
        self.enc_token_id, self.pad_token_id, self.cls_token_id, self.sep_token_id = ENC_ID, PAD_ID, CLS_ID, SEP_ID

    def _pid(self, piece):
        if piece not in self._piece2id:
            i = 1000 + zlib.crc32(piece.encode()) % 28000
            while i in self._id2piece:
                i += 1
            self._piece2id[piece] = i
            self._id2piece[i] = piece
        return self._piece2id[piece]

    def pieces(self, word):
        word = word.lower()
        if len(word) <= self.split_len:
            return [word]
        out = [word[:self.piece_len]]
        rest = word[self.piece_len:]
        while rest:
            out.append("##" + rest[:self.piece_len])
            rest = rest[self.piece_len:]
        return out

    def encode(self, text):
        ids = [CLS_ID]
        for w in text.split():
            ids.extend(self._pid(p) for p in self.pieces(w))
        ids.append(SEP_ID)
        return ids

    def decode(self, ids):
        return " ".join(self._id2piece.get(int(i), "[UNK]") for i in ids)

    def __call__(self, text, padding="longest", max_length=None, truncation=False, return_tensors="pt"):
        if isinstance(text, str):
            text = [text]
        enc = [self.encode(t) for t in text]
        if truncation and max_length:
            enc = [e[:max_length] for e in enc]
        L = max_length if padding == "max_length" else max(len(e) for e in enc)
        ids = torch.full((len(enc), L), PAD_ID, dtype=torch.long)
        att = torch.zeros((len(enc), L), dtype=torch.long)
        for i, e in enumerate(enc):
            ids[i, :len(e)] = torch.tensor(e)
            att[i, :len(e)] = 1
        return BatchEncoding(ids, att)


class BatchEncoding:
    def __init__(self, input_ids, attention_mask):
        self.input_ids, self.attention_mask = input_ids, attention_mask

    def to(self, device):
        return BatchEncoding(self.input_ids.to(device), self.attention_mask.to(device))


def saliency_maps(seed, C, P, zero_frac=0.5):
    """relu(normal)^2-like maps with ~zero_frac exact zeros (mimics clamp(grad, 0) sparsity)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(C, P, P, generator=g)
    x = torch.relu(x - torch.quantile(x.flatten(), zero_frac)) ** 2
    return x.contiguous()


def guide_image(seed, H, W, kind="natural"):
    """uint8 [H,W,3] CRF guide image: 'natural' = low-frequency cosines + noise, 'noise' = iid uniform."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = np.zeros((H, W, 3))
    for c in range(3):
        acc = np.zeros((H, W))
        for _ in range(6):
            fx, fy = rng.uniform(0.2, 3.0, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.cos(2 * np.pi * (fx * xx / W + fy * yy / H) + ph)
        img[:, :, c] = 128 + 60 * acc / 6 * 2.0 + rng.normal(0, 4, (H, W))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def gt_labels(seed, H, W, n_class, ignore_frac=0.02, blocky=True):
    """float32 [H,W] ground truth in [0,n_class) with ~ignore_frac pixels = 255."""
    rng = np.random.default_rng(seed)
    if blocky:
        small = rng.integers(0, n_class, ((H + 15) // 16, (W + 15) // 16))
        lab = np.kron(small, np.ones((16, 16), dtype=np.int64))[:H, :W]
    else:
        lab = rng.integers(0, n_class, (H, W))
    lab = lab.astype(np.float32)
    lab[rng.random((H, W)) < ignore_frac] = 255.0
    return lab


class SynthGradcamFn:
    """Deterministic stand-in for compute_gradcam_ensemble(...)[layer][head]: a [B,T-1,P,P] map that depends
    on the (possibly patch-zeroed) image, so the Salience DropOut loop has something to react to."""

    def __init__(self, seed, B, T, P):
        g = torch.Generator().manual_seed(seed)
        self.w = torch.rand(B, T - 1, 1, 1, generator=g) + 0.1
        self.noise = torch.rand(B, T - 1, P, P, generator=g) * 0.05
        self.P = P
        self.calls = 0

    def __call__(self, imgs, attention_rows=None):
        B, _, S, _ = imgs.shape
        P = self.P
        cell = imgs.reshape(B, 3, P, S // P, P, S // P).abs().mean(dim=(1, 3, 5))  # [B,P,P]
        g = torch.relu(cell.unsqueeze(1) * self.w + self.noise - 0.3 - 0.02 * self.calls)
        self.calls += 1
        if attention_rows is not None:  # rows past a caption's length are zeroed by the mask (BITM:427)
            g = g * attention_rows.view(B, -1, 1, 1)
        return g
