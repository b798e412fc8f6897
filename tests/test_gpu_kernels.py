"""GPU parity, kernel level: every C-ABI entry point against the oracle on the same
seeded inputs (bit-exact for ids / hashes, fp32 tolerance stated per test)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import criteo, farmhash, models as om, synth, tfsem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ops():
    from recsys_b200 import ops
    return ops


def _criteo_layout(D=16, full=False):
    from recsys_b200 import feature_column as fc
    from recsys_b200.fm import fm
    lin, emb = fm.build_feature_columns(D, full_cardinality=full)
    return fc.layout(emb), lin


def _to_torch_features(feats):
    out = {}
    for k, v in feats.items():
        out[k] = torch.from_numpy(np.asarray(v)) if np.asarray(v).dtype.kind in "fiu" else v
    return out


# ------------------------------------------------------------------ id pipeline
@pytest.mark.parametrize("B", [1, 7, 256, 4099])
def test_criteo_rows_bit_exact(cuda, B):
    ops = _ops()
    lay, _ = _criteo_layout()
    spec = criteo.CriteoSpec()
    feats, _ = criteo.synthetic_features(B, seed=B, spec=spec, dist="zipf")
    want = criteo.criteo_rows(feats, spec)
    pipe = ops.IdPipeline(lay, cuda)
    rows, logx = pipe(_to_torch_features(feats), want_logx=True)
    assert rows.dtype == torch.int32 and tuple(rows.shape) == (B, 39)
    got = rows.cpu().numpy().astype(np.int64)
    # logf on the device vs numpy may differ by 1 ulp: a mismatch is only legal when the
    # log value sits within 1 ulp of a boundary.
    bad = np.argwhere(got != want)
    lx = criteo.criteo_logx(feats, spec)
    for b, f in bad:
        key = spec.fields[f]
        assert key in criteo.CONT
        v = lx[b, spec.cont_fields.index(key)]
        bnd = np.asarray(spec.boundaries(key), np.float32)
        assert np.min(np.abs(bnd - v)) <= 2 * np.spacing(np.float32(abs(v) + 1e-30))
    assert len(bad) <= max(1, B // 1000)
    assert np.allclose(logx.cpu().numpy(), lx, rtol=3e-7, atol=1e-7, equal_nan=True)
    assert int(pipe.status.item()) == 0
    # the few-CTA variant for a copy stream (ctr_criteo_rows_bg): the same ids, bit for bit
    rows_bg = pipe(_to_torch_features(feats), background=4)
    assert torch.equal(rows_bg, rows)


@pytest.mark.parametrize("B", [1, 7, 256, 4099])
@pytest.mark.parametrize("record", [True, False])
def test_fused_ids_and_lookup_equal_the_two_launch_path(cuda, B, record):
    """ctr_embed_fwd_raw (id pipeline as the first stage of the lookup kernel) against
    ctr_criteo_rows + ctr_embed_fwd: identical ids, logx and lookup outputs, bit for bit."""
    ops = _ops()
    lay, _ = _criteo_layout()
    spec = criteo.CriteoSpec()
    feats, _ = criteo.synthetic_features(B, seed=B + 1, spec=spec, dist="zipf")
    tf = _to_torch_features(feats)
    pipe = ops.IdPipeline(lay, cuda)
    emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=(1 << 39) - 1, record=record)
    with torch.no_grad():
        rows_a, logx_a = pipe(tf, want_logx=True)
        Ea, y1a, y2a, _ = emb.lookup(rows_a, want_lo=True)
        Elo_a = emb.last_E_lo
        rows_b, logx_b, Eb, y1b, y2b, _ = emb.lookup_features(pipe, tf, want_logx=True, want_lo=True)
        Elo_b = emb.last_E_lo
    torch.cuda.synchronize()
    assert torch.equal(rows_a, rows_b) and torch.equal(logx_a, logx_b)
    assert torch.equal(Ea, Eb) and torch.equal(y1a, y1b) and torch.equal(y2a, y2b)
    assert torch.equal(Elo_a, Elo_b)
    assert int(pipe.status.item()) == 0


def test_criteo_rows_real_shard_strings(cuda):
    """Raw byte strings (b'NULL' default included) hashed on the device == oracle == fixture."""
    ops = _ops()
    lay, _ = _criteo_layout()
    z = np.load(os.path.join(GOLD, "criteo_shard256.npz"))
    feats = {k: torch.from_numpy(z[k]) for k in criteo.CONT}
    feats.update({k: np.array([bytes(v) for v in z[k]], dtype=object).reshape(-1, 1)
                  for k in criteo.CAT})
    rows = ops.IdPipeline(lay, cuda)(feats)
    assert np.array_equal(rows.cpu().numpy().astype(np.int64), z["rows"])


def test_hash_strings_all_length_branches(cuda):
    ops = _ops()
    rng = np.random.default_rng(0)
    strs = [b"Hello", b"TensorFlow", b"2.x", b"", b"NULL"]
    for n in list(range(1, 70)) + [100, 127, 128, 129, 200, 333]:
        strs.append(bytes(rng.integers(0, 256, size=n, dtype=np.uint8)))
    for nb in (3, 100000, 2 ** 31 - 1):
        got = ops.hash_strings(strs, nb, cuda).cpu().numpy()
        want = np.array([farmhash.hash_bucket_fast(s, nb) for s in strs])
        assert np.array_equal(got, want)
    assert ops.hash_strings(strs[:3], 3, cuda).cpu().tolist() == [0, 2, 2]   # TF doc example


def test_out_of_range_categorical_sets_status(cuda):
    ops = _ops()
    lay, _ = _criteo_layout()
    spec = criteo.CriteoSpec()
    feats, _ = criteo.synthetic_features(8, seed=1, spec=spec)
    feats["_c14"] = feats["_c14"] + 10 ** 7
    pipe = ops.IdPipeline(lay, cuda)
    rows = pipe(_to_torch_features(feats))
    f = spec.fields.index("_c14")
    r = rows.cpu().numpy()[:, f]
    assert ((r >= spec.offsets[f]) & (r < spec.offsets[f + 1])).all()
    assert int(pipe.status.item()) == 1


# ------------------------------------------------------------------ embed fwd/bwd
def _rand_rows(B, offsets, seed, hot=False):
    rng = np.random.default_rng(seed)
    cols = []
    for f in range(len(offsets) - 1):
        n = offsets[f + 1] - offsets[f]
        ids = np.zeros(B, np.int64) if (hot and f % 3 == 0) else rng.integers(0, n, size=B)
        cols.append(ids + offsets[f])
    return np.stack(cols, 1)


@pytest.mark.parametrize("B,F,D", [(1, 39, 16), (5, 39, 16), (16, 39, 16), (4096, 39, 16),
                                   (333, 2, 32), (130, 39, 32), (257, 64, 8), (64, 7, 16)])
@pytest.mark.parametrize("record", [True, False])
def test_embed_fwd_matches_oracle(cuda, B, F, D, record):
    ops = _ops()
    from recsys_b200 import feature_column as fc
    rng = np.random.default_rng(B + F + D)
    nrows = [int(n) for n in rng.integers(3, 400, size=F)]
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("k%02d" % i, n), D)
            for i, n in enumerate(nrows)]
    lay = fc.layout(cols)
    mask = (int.from_bytes(rng.bytes(8), "little") & ((1 << F) - 1)) | 1
    emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=mask, record=record)
    rows_np = _rand_rows(B, lay.offsets, seed=B, hot=True)
    rows = torch.from_numpy(rows_np).to(cuda, torch.int32)
    with torch.no_grad():
        E, y1, y2, _ = emb.lookup(rows)
    t64 = emb.table.double().cpu()
    Eo = t64[torch.from_numpy(rows_np)]                               # [B,F,D]
    assert torch.equal(E.cpu(), Eo.float().reshape(B, F * D))        # gather is a copy: bit exact
    fm_mask = torch.tensor([(mask >> f) & 1 for f in range(F)], dtype=torch.float64)
    y1o = (emb.w1.double().cpu()[torch.from_numpy(rows_np)] * fm_mask).sum(1)
    y2o = om.fm_second_order(Eo).reshape(-1)
    assert torch.allclose(y1.cpu().double(), y1o, rtol=1e-5, atol=1e-6)
    assert torch.allclose(y2.cpu().double(), y2o, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,N,fields", [(4096, 100, "criteo"), (300, 100, "criteo"), (256, 128, "criteo"),
                                        (1000, 16, "criteo"), (515, 64, 12), (2048, 100, 40)])
def test_embed_tower_fwd_matches_unfused(cuda, B, N, fields):
    """ctr_embed_tower_fwd (ids + gather + FM terms + first tower layer on tcgen05, one launch)
    against ctr_embed_fwd_raw + float64 relu(E . W0 + b0): ids, E and E_lo bit-identical; S / y1 /
    y2 to summation-order tolerance; act0 to 3xTF32 tolerance (2e-5 of the row scale); the
    per-cluster column sums add up to the column sums of act0."""
    ops = _ops()
    from recsys_b200 import _lib
    from recsys_b200 import feature_column as fc
    lib = _lib.load()
    if fields == "criteo":
        lay, _ = _criteo_layout()
        spec = criteo.CriteoSpec()
        feats, _ = criteo.synthetic_features(B, seed=B + N, spec=spec, dist="zipf")
        tf = _to_torch_features(feats)
    else:
        rng = np.random.default_rng(fields)
        nrows = [int(n) for n in rng.integers(2, 5000, size=fields)]
        cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("k%02d" % i, n), 16)
                for i, n in enumerate(nrows)]
        lay = fc.layout(cols)
        tf = {"k%02d" % i: torch.from_numpy(rng.integers(0, n, size=B)) for i, n in enumerate(nrows)}
    F, D = lay.F, 16
    pipe = ops.IdPipeline(lay, cuda)
    emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=(1 << F) - 1 - 4, seed=3)
    g = torch.Generator(device="cpu").manual_seed(N)
    W0 = (torch.randn(F * D, N, generator=g) * (2.0 / (F * D)) ** 0.5).to(cuda)
    b0 = (torch.randn(N, generator=g) * 0.1).to(cuda)
    W0_lo = ops.split_lo(W0)
    with torch.no_grad():
        rows_a, _, Ea, y1a, y2a, _ = emb.lookup_features(pipe, tf, want_lo=True)
        Elo_a = emb.last_E_lo.clone()
        Sa = Ea.view(B, F, D).double().sum(1)
    cont, cat = pipe.pack(tf)
    p = ops._p
    rows = torch.full((B, F), -7, dtype=torch.int32, device=cuda)
    E = torch.full((B, F * D), float("nan"), device=cuda)
    E_lo = torch.full_like(E, float("nan"))
    S = torch.empty(B, D, device=cuda)
    y1 = torch.empty(B, device=cuda)
    y2 = torch.empty(B, device=cuda)
    act0 = torch.full((B, N), float("nan"), device=cuda)
    nparts = (B + 127) // 128
    parts = torch.full((nparts, 2, N), float("nan"), device=cuda)
    zbuf = torch.ones(64, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.ctr_embed_tower_fwd(p(emb.table), p(emb.w1), p(cont), len(pipe.cont_keys), p(cat),
                                 len(pipe.cat_keys), p(pipe.fields_dev), p(pipe.bnd_dev), pipe.n_bnd, None, p(rows),
                                 p(pipe.status), B, F, D, emb.w1_fields, p(E), p(E_lo), p(S), p(y1), p(y2),
                                 emb.ld, emb.ld1, p(W0), p(W0_lo), p(b0), N, p(act0), p(parts), p(zbuf),
                                 zbuf.numel(), st)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert torch.equal(rows, rows_a)
    assert torch.equal(E, Ea) and torch.equal(E_lo, Elo_a)
    assert float(zbuf.abs().max()) == 0.0
    assert torch.allclose(S.double().cpu(), Sa.cpu(), rtol=1e-5, atol=1e-5)
    assert torch.allclose(y1, y1a, rtol=1e-5, atol=1e-6)
    assert torch.allclose(y2, y2a, rtol=1e-4, atol=1e-4)
    want = torch.relu(Ea.double() @ W0.double() + b0.double())
    scale = float(want.abs().max())
    err = float((act0.double() - want).abs().max())
    assert err <= 2e-5 * scale + 1e-6, "act0 differs: %g (scale %g)" % (err, scale)
    cs = torch.stack([act0.double().sum(0), (act0.double() ** 2).sum(0)])
    got = parts.double().sum(0)
    assert torch.allclose(got, cs, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("B,N,F,use_fm", [(4096, 100, 39, True), (300, 100, 39, True), (256, 128, 39, False),
                                          (515, 64, 12, True), (1000, 16, 40, True)])
def test_tower_embed_bwd_matches_unfused(cuda, B, N, F, use_fm):
    """ctr_tower_embed_bwd (dE = dpre0 . W0^T in TMEM, scattered from there) against ctr_embed_bwd fed
    with the float64 product: table / first-order gradient accumulators to 3xTF32 tolerance.
    Layouts with one-row, <= 32-row and large fields; ragged B."""
    ops = _ops()
    from recsys_b200 import _lib
    from recsys_b200 import feature_column as fc
    lib = _lib.load()
    D = 16
    rng = np.random.default_rng(B + N + F)
    nrows = [int(n) for n in rng.integers(2, 4000, size=F)]
    nrows[0], nrows[1], nrows[F // 2] = 1, 7, 32
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("k%02d" % i, n), D)
            for i, n in enumerate(nrows)]
    lay = fc.layout(cols)
    mask = (1 << F) - 1 - 2
    rows_np = _rand_rows(B, lay.offsets, seed=B)
    rows = torch.from_numpy(rows_np).to(cuda, torch.int32)
    g = torch.Generator(device="cpu").manual_seed(N)
    W0 = (torch.randn(F * D, N, generator=g) * (2.0 / (F * D)) ** 0.5).to(cuda)
    dpre0 = (torch.randn(B, N, generator=g) * 1e-3).to(cuda)
    dy1 = torch.randn(B, generator=g).to(cuda) * 1e-3
    dy2 = torch.randn(B, generator=g).to(cuda) * 1e-3 if use_fm else None
    res = []
    for fused in (False, True):
        emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=mask, seed=2)
        with torch.no_grad():
            E, _, _, _ = emb.lookup(rows)
        S = E.view(B, F, D).sum(1).contiguous()
        p = ops._p
        st = torch.cuda.current_stream().cuda_stream
        if fused:
            dpre0_lo, W0_lo = ops.split_lo(dpre0), ops.split_lo(W0)
            rc = lib.ctr_tower_embed_bwd(p(dpre0), p(dpre0_lo), p(W0), p(W0_lo), N,
                                         p(rows), p(E), p(S), p(dy2), p(dy1), mask, emb._offsets_host, B, F,
                                         D, p(emb.dtable), p(emb.dw1), emb.ld, emb.ld1, st)
        else:
            dE = (dpre0.double() @ W0.double().t()).float().contiguous()
            rc = lib.ctr_embed_bwd(p(rows), p(dE), p(E), p(emb.table), p(S), p(dy2), p(dy1), mask,
                                   emb._offsets_host, B, F, D, p(emb.dtable), p(emb.dw1), emb.ld, emb.ld1,
                                   st)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        res.append((emb.dtable.clone(), emb.dw1.clone()))
    (gt, gw), (ft, fw) = res
    scale = float(gt.abs().max())
    assert scale > 0
    assert float((ft - gt).abs().max()) <= 2e-5 * scale + 1e-9
    assert float((fw - gw).abs().max()) <= 1e-5 * float(gw.abs().max()) + 1e-9


@pytest.mark.parametrize("B,D,use_dE,use_fm,regather", [
    (1, 16, True, True, False), (77, 16, True, True, False), (4096, 16, True, True, False),
    (300, 16, False, True, True), (300, 16, True, False, False), (515, 32, True, True, False),
    (129, 8, True, True, True)])
@pytest.mark.parametrize("record", [True, False])
def test_embed_bwd_matches_oracle(cuda, B, D, use_dE, use_fm, regather, record):
    """dtable / dw1 against the dense autograd gradient of the oracle; includes fields with
    3..32 rows (register one-hot path), a field where every sample hits one row, and big fields."""
    ops = _ops()
    from recsys_b200 import _lib
    from recsys_b200 import feature_column as fc
    nrows = [3, 10, 32, 33, 7, 1000, 50000, 4, 64, 31, 200]
    F = len(nrows)
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("k%02d" % i, n), D)
            for i, n in enumerate(nrows)]
    lay = fc.layout(cols)
    mask = 0b10110101101
    emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=mask, record=record)
    rows_np = _rand_rows(B, lay.offsets, seed=B + D, hot=True)
    rows = torch.from_numpy(rows_np).to(cuda, torch.int32)
    g = torch.Generator().manual_seed(B)
    dE = torch.randn(B, F * D, generator=g)
    dy1 = torch.randn(B, generator=g)
    dy2 = torch.randn(B, generator=g)
    with torch.no_grad():
        E, y1, y2, _ = emb.lookup(rows)
    S = E.view(B, F, D).sum(1).contiguous()
    lib = _lib.load()
    offs = (C.c_int64 * (F + 1))(*lay.offsets)
    p = ops._p
    dEc, dy2c, dy1c = dE.to(cuda), dy2.to(cuda), dy1.to(cuda)
    rc = lib.ctr_embed_bwd(p(rows), p(dEc) if use_dE else None, None if regather else p(E),
                           p(emb.table), p(S) if use_fm else None, p(dy2c) if use_fm else None,
                           p(dy1c), mask, offs, B, F, D, p(emb.dtable), p(emb.dw1), emb.ld, emb.ld1,
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    # oracle: autograd through gather + FM second order
    t = emb.table.double().cpu().requires_grad_(True)
    w = emb.w1.double().cpu().requires_grad_(True)
    r = torch.from_numpy(rows_np)
    Eo = t[r]
    loss = torch.zeros((), dtype=torch.float64)
    if use_dE:
        loss = loss + (Eo.reshape(B, -1) * dE.double()).sum()
    if use_fm:
        loss = loss + (om.fm_second_order(Eo).reshape(-1) * dy2.double()).sum()
    fm_mask = torch.tensor([(mask >> f) & 1 for f in range(F)], dtype=torch.float64)
    loss = loss + ((w[r] * fm_mask).sum(1) * dy1.double()).sum()
    loss.backward()
    scale = float(t.grad.abs().max()) + 1e-12
    assert float((emb.dtable.cpu().double() - t.grad).abs().max()) <= 2e-5 * scale + 1e-5
    assert float((emb.dw1.cpu().double() - w.grad).abs().max()) <= 1e-5 * float(w.grad.abs().max()) + 1e-5


@pytest.mark.parametrize("B,D,use_dE,use_fm", [(1, 16, True, True), (77, 16, True, True),
                                               (4096, 16, True, True), (300, 16, False, True),
                                               (300, 16, True, False), (515, 32, True, True),
                                               (129, 8, True, True)])
def test_fused_scatter_adam_matches_tf_adam_on_the_oracle_gradient(cuda, B, D, use_dE, use_fm):
    """ctr_count_rows + ctr_embed_bwd_adam (scatter-add and row optimiser in one pass): three
    steps against the dense autograd gradient of the oracle fed to tfsem.TFAdam (lazy rows).
    Fields with 3..32 rows (shared-memory tile), a field where every sample hits one row,
    in-warp duplicates (match-any aggregation) and big fields; afterwards the records'
    accumulators and lookup counts are all zero again."""
    ops = _ops()
    from recsys_b200 import _lib
    from recsys_b200 import feature_column as fc
    nrows = [3, 10, 32, 33, 7, 1000, 50000, 4, 64, 31, 200]
    F = len(nrows)
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("k%02d" % i, n), D)
            for i, n in enumerate(nrows)]
    lay = fc.layout(cols)
    mask = 0b10110101101
    emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=mask, adam_mode="lazy", seed=2)
    assert emb.record and emb.can_fuse
    p64 = {"emb": emb.table.double().cpu().clone(), "w1": emb.w1.double().cpu().clone()}
    opt = tfsem.TFAdam(p64, lr=1e-2)
    st = ops.TFAdamState(lr=1e-2, device=cuda)
    lib = _lib.load()
    offs = (C.c_int64 * (F + 1))(*lay.offsets)
    P = ops._p
    fm_mask = torch.tensor([(mask >> f) & 1 for f in range(F)], dtype=torch.float64)
    for step in range(3):
        rows_np = _rand_rows(B, lay.offsets, seed=B + D + step, hot=True)
        rows = torch.from_numpy(rows_np).to(cuda, torch.int32)
        g = torch.Generator().manual_seed(B + step)
        dE = torch.randn(B, F * D, generator=g)
        dy1 = torch.randn(B, generator=g)
        dy2 = torch.randn(B, generator=g)
        with torch.no_grad():
            E, _, _, _ = emb.lookup(rows)
        S = E.view(B, F, D).sum(1).contiguous()
        dEc, dy2c, dy1c = dE.to(cuda), dy2.to(cuda), dy1.to(cuda)
        # oracle gradient (autograd through gather + FM second order), then the TF Adam rule
        t = p64["emb"].clone().requires_grad_(True)
        w = p64["w1"].clone().requires_grad_(True)
        r = torch.from_numpy(rows_np)
        Eo = t[r]
        loss = ((w[r] * fm_mask).sum(1) * dy1.double()).sum()
        if use_dE:
            loss = loss + (Eo.reshape(B, -1) * dE.double()).sum()
        if use_fm:
            loss = loss + (om.fm_second_order(Eo).reshape(-1) * dy2.double()).sum()
        loss.backward()
        gw = w.grad
        # first-order weights exist only for the masked fields: lazy rows = their lookups
        # (the kernel never touches theta1 of a row whose field has no first-order term)
        fsel = [f for f in range(F) if (mask >> f) & 1]
        opt.step(p64, {"emb": t.grad, "w1": gw},
                 lazy_rows={"emb": r.reshape(-1), "w1": r[:, fsel].reshape(-1)})
        lr_t = st.next_lr_t()
        sm = torch.cuda.current_stream().cuda_stream
        assert lib.ctr_count_rows(P(rows), rows.numel(), D, P(emb.rec), emb.ld, sm) == 0
        rc = lib.ctr_embed_bwd_adam(P(rows), P(dEc) if use_dE else None, P(S) if use_fm else None,
                                    P(dy2c) if use_fm else None, P(dy1c), mask, offs, B, F, D,
                                    P(emb.rec), emb.ld, lr_t, st.beta1, st.beta2, st.eps,
                                    st.state_ptr, sm)
        assert rc == 0, _lib.last_error()
        st.advance()
        torch.cuda.synchronize()
        assert float(emb.dtable.abs().max()) == 0.0 and float(emb.dw1.abs().max()) == 0.0
        assert int(emb.rec[:, 4 * D + 5].view(torch.int32).abs().max()) == 0      # cnt
        assert float(emb.rec[:, 4 * D + 6].abs().max()) == 0.0                    # c
        assert int(emb.rec[:, 4 * D + 7].view(torch.int32).abs().max()) == 0      # arr
    # w1 of unmasked fields: the oracle's lazy rule decays nothing there either (zero gradient,
    # rows not listed) - compare everything
    assert torch.allclose(emb.table.cpu().double(), p64["emb"], rtol=1e-5, atol=2e-6)
    sel = torch.zeros(lay.total_rows, dtype=torch.bool)
    for f in range(F):
        if (mask >> f) & 1:
            sel[lay.offsets[f]:lay.offsets[f + 1]] = True
    assert torch.allclose(emb.w1.cpu().double()[sel], p64["w1"][sel], rtol=1e-5, atol=2e-6)


def test_embed_rejects_bad_arguments(cuda):
    from recsys_b200 import _lib
    lib = _lib.load()
    t = torch.zeros(10, 16, device=cuda)
    rows = torch.zeros(4, 3, dtype=torch.int32, device=cuda)
    rc = lib.ctr_embed_fwd(t.data_ptr(), None, rows.data_ptr(), 4, 3, 12, 0, None, None, None, None,
                           None, None, 0, None, None, 0, 0, None)
    assert rc == -1 and "D must be" in _lib.last_error()
    rc = lib.ctr_embed_fwd(t.data_ptr(), None, rows.data_ptr(), 4, 65, 16, 0, None, None, None, None,
                           None, None, 0, None, None, 0, 0, None)
    assert rc == -1
    rc = lib.ctr_embed_fwd(t.data_ptr(), None, rows.data_ptr(), 0, 3, 16, 0, None, None, None, None,
                           None, None, 0, None, None, 0, 0, None)
    assert rc == 0                                                 # empty batch is a no-op


# ------------------------------------------------------------------- DCN cross
@pytest.mark.parametrize("B,W,L", [(1, 624, 4), (100, 624, 4), (4096, 624, 4), (33, 64, 1),
                                   (257, 1248, 3), (65, 128, 6)])
def test_dcn_cross_fwd_bwd(cuda, B, W, L):
    ops = _ops()
    g = torch.Generator().manual_seed(B + W)
    x0 = torch.randn(B, W, generator=g) * 0.5
    w = torch.randn(L, W, generator=g) * 0.05
    b = torch.randn(L, W, generator=g) * 0.05
    dxl = torch.randn(B, W, generator=g)
    x0c = x0.to(cuda).requires_grad_(True)
    wc = w.to(cuda).requires_grad_(True)
    bc = b.to(cuda).requires_grad_(True)
    xl = ops.dcn_cross(x0c, wc, bc)
    xl.backward(dxl.to(cuda))
    p = {}
    x0o = x0.double().requires_grad_(True)
    for l in range(L):
        p[f"cross.{l}.w"] = w[l].double().requires_grad_(True)
        p[f"cross.{l}.b"] = b[l].double().requires_grad_(True)
    xlo = om.dcn_cross(p, x0o)
    xlo.backward(dxl.double())
    assert torch.allclose(xl.detach().cpu().double(), xlo.detach(), rtol=1e-4, atol=1e-5)
    s = float(x0o.grad.abs().max())
    assert float((x0c.grad.cpu().double() - x0o.grad).abs().max()) <= 1e-4 * s + 1e-6
    for l in range(L):
        gw, gb = p[f"cross.{l}.w"].grad, p[f"cross.{l}.b"].grad
        assert float((wc.grad[l].cpu().double() - gw).abs().max()) <= 1e-4 * float(gw.abs().max()) + 1e-5
        assert float((bc.grad[l].cpu().double() - gb).abs().max()) <= 1e-4 * float(gb.abs().max()) + 1e-5


# ------------------------------------------------------------------------- Adam
def test_adam_rows_and_dense_match_tf_rule(cuda):
    ops = _ops()
    from recsys_b200 import feature_column as fc
    D = 16
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("a", 50), D),
            fc.embedding_column(fc.categorical_column_with_hash_bucket("b", 5000), D)]
    lay = fc.layout(cols)
    for mode in ("lazy", "lazy-planar", "exact_tf"):      # row records / planar arrays / dense apply
        record = {"lazy": True, "lazy-planar": False, "exact_tf": None}[mode]
        mode = mode.split("-")[0]
        emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=0b11, adam_mode=mode, seed=1,
                                 record=record)
        assert emb.record == (record is True)
        p = {"emb": emb.table.double().cpu().clone(), "w1": emb.w1.double().cpu().clone()}
        opt = tfsem.TFAdam(p, lr=1e-2)
        st = ops.TFAdamState(lr=1e-2, device=cuda if mode == "lazy" else None)
        rng = np.random.default_rng(0)
        for step in range(3):
            rows_np = _rand_rows(300, lay.offsets, seed=step)
            rows = torch.from_numpy(rows_np).to(cuda, torch.int32)
            g = torch.from_numpy(rng.normal(size=(lay.total_rows, D))).float()
            g1 = torch.from_numpy(rng.normal(size=lay.total_rows)).float()
            touched = torch.zeros(lay.total_rows, dtype=torch.bool)
            touched[torch.from_numpy(rows_np).reshape(-1)] = True
            g[~touched] = 0
            g1[~touched] = 0
            emb.dtable.copy_(g)
            emb.dw1.copy_(g1)
            emb.adam_step(rows, st.next_lr_t(), st)
            st.advance()        # no dense parameters here: the caller moves the device schedule on
            lazy = {"emb": torch.from_numpy(rows_np).reshape(-1),
                    "w1": torch.from_numpy(rows_np).reshape(-1)} if mode == "lazy" else None
            opt.step(p, {"emb": g.double(), "w1": g1.double()}, lazy_rows=lazy)
            assert float(emb.dtable.abs().max()) == 0.0 and float(emb.dw1.abs().max()) == 0.0
        assert torch.allclose(emb.table.cpu().double(), p["emb"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(emb.w1.cpu().double(), p["w1"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("D", [8, 16, 32])
def test_adam_rows_bf_equals_flat_kernel(cuda, D, monkeypatch):
    """ctr_adam_rows_bf (field-major warps, in-warp de-duplication, winners only) against
    ctr_adam_rows on the same [B, F] ids, both in-flight depths: bit-identical tables, moments and
    cleared gradients; ragged B, a one-row field, negative (padding) ids."""
    ops = _ops()
    from recsys_b200 import _lib
    from recsys_b200 import feature_column as fc
    lib = _lib.load()
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("a", 1), D),
            fc.embedding_column(fc.categorical_column_with_hash_bucket("b", 7), D),
            fc.embedding_column(fc.categorical_column_with_hash_bucket("c", 3000), D),
            fc.embedding_column(fc.categorical_column_with_hash_bucket("d", 100000), D)]
    lay = fc.layout(cols)
    B = 1000 + 13
    rows_np = _rand_rows(B, lay.offsets, seed=5)
    rows_np[::17, 2] = -1
    rows = torch.from_numpy(rows_np).to(cuda, torch.int32)
    g = torch.randn(lay.total_rows, D, device=cuda)
    g1 = torch.randn(lay.total_rows, device=cuda)
    results = []
    for variant in ("flat", "n1", "n2"):
        monkeypatch.setenv("CTR_ADAM_ROWS_BF", "0" if variant == "flat" else "1")
        assert lib.ctr_set_option(b"adam_rows_inflight", int(variant[1]) if variant != "flat" else 1) == 0
        emb = ops.FieldEmbedding(lay, cuda, with_w1=True, w1_fields=0b1111, adam_mode="lazy", seed=1)
        st = ops.TFAdamState(lr=1e-2, device=cuda)
        for step in range(2):
            emb.dtable.copy_(g * (step + 1))
            emb.dw1.copy_(g1)
            emb.adam_step(rows, st.next_lr_t(), st)
            st.advance()
        torch.cuda.synchronize()
        results.append(emb.rec.clone())
    assert lib.ctr_set_option(b"adam_rows_inflight", 1) == 0
    valid = torch.from_numpy(rows_np[rows_np >= 0].reshape(-1)).long()
    touched = torch.zeros(lay.total_rows, dtype=torch.bool)
    touched[valid] = True
    ref = results[0].cpu()
    assert float(ref[touched][:, 3 * D:4 * D].abs().max()) == 0.0          # g cleared where touched
    assert float((ref[~touched][:, 3 * D:4 * D] - (2 * g).cpu()[~touched]).abs().max()) == 0.0
    for r in results[1:]:
        assert torch.equal(r.cpu(), ref)


def test_adam_schedule_counts_past_2_to_the_24(cuda):
    """the device step counter must not saturate (a float32 counter stops at 2^24,
    after which the lazy optimiser's claim tag repeats and rows stop updating).  It is a uint32
    bit pattern: two steps from t = 2^24 still update the touched rows and end at 2^24 + 2."""
    ops = _ops()
    from recsys_b200 import feature_column as fc
    D = 16
    lay = fc.layout([fc.embedding_column(fc.categorical_column_with_hash_bucket("a", 100), D)])
    emb = ops.FieldEmbedding(lay, cuda, with_w1=False, adam_mode="lazy", seed=1)
    st = ops.TFAdamState(lr=1e-2, device=cuda)
    st.state.view(torch.int32)[0] = 2 ** 24 - 1
    st.advance()
    assert int(st.state.view(torch.int32)[0]) == 2 ** 24
    rows = torch.arange(10, device=cuda, dtype=torch.int32).reshape(-1, 1)
    prev = emb.table.clone()
    for step in range(2):
        emb.dtable[:10] = 1.0
        emb.adam_step(rows, st.next_lr_t(), st)
        st.advance()
        torch.cuda.synchronize()
        moved = (emb.table[:10] - prev[:10]).abs().min()
        assert float(moved) > 1e-4, "step %d: the rows did not move" % step
        assert float(emb.dtable.abs().max()) == 0.0
        prev = emb.table.clone()
    assert int(st.state.view(torch.int32)[0]) == 2 ** 24 + 2


# -------------------------------------------------------------------------- CIN
def _cin_oracle(E, Ws, bs, m, D):
    p = {}
    for k, (W, b) in enumerate(zip(Ws, bs)):
        p[f"cin.{k}.w"], p[f"cin.{k}.b"] = W, b
    B = E.shape[0]
    X0 = E.view(B, m, D)
    hidden, finals = X0, []
    for k in range(len(Ws)):
        Hp = hidden.shape[1]
        z = torch.einsum("bid,bjd->bdij", X0, hidden).reshape(B, D, m * Hp)
        out = torch.relu(z @ Ws[k] + bs[k]).permute(0, 2, 1)
        finals.append(out)
        hidden = out
    return torch.cat(finals, 1).sum(-1)


CIN_TOL = {"fp32": 2e-5, "tf32x3": 2e-5, "tf32": 3e-3}


def _cin_layer_case(B, D, m, Hp, H, seed):
    g = torch.Generator().manual_seed(seed)
    M = B * D
    ld0 = (m + 3) // 4 * 4
    X0t = torch.zeros(M, ld0)
    X0t[:, :m] = torch.randn(M, m, generator=g) * 0.3
    Xp = torch.randn(M, Hp, generator=g) * 0.3
    W = torch.randn(m * Hp, H, generator=g) * (2.0 / (m * Hp + H)) ** 0.5
    bias = torch.randn(H, generator=g) * 0.05
    dpre = torch.randn(M, H, generator=g)
    return X0t, ld0, Xp, W, bias, dpre


@pytest.mark.parametrize("prec", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("B,D,m,Hp,H", [(8, 16, 39, 39, 128), (8, 16, 39, 128, 128), (19, 16, 39, 128, 128),
                                        (5, 16, 39, 20, 10), (5, 16, 39, 10, 10), (3, 8, 5, 5, 16)])
def test_cin_layer_abi(cuda, prec, B, D, m, Hp, H):
    """ctr_cin_layer_fwd / _bwd against the einsum restatement of xdeepfm/xdeepfm.py:145-169,
    with the pre-activation gradient injected (no ReLU-mask sensitivity).  Tolerances relative
    to the max-norm: fp32 / 3xTF32 2e-5, plain TF32 (rounded operands) 3e-3."""
    ops = _ops()
    from recsys_b200 import _lib
    lib = _lib.load()
    X0t, ld0, Xp, W, bias, dpre = _cin_layer_case(B, D, m, Hp, H, seed=B + Hp + H)
    if Hp == m:
        Xp = X0t[:, :m].clone()
    M = B * D
    pc = ops.CIN_PREC[prec]
    dev = lambda t: t.to(cuda).contiguous()
    X0c, Xpc, Wc, bc, dc = dev(X0t), dev(Xp), dev(W), dev(bias), dev(dpre)
    nbytes = int(lib.ctr_cin_workspace_bytes(B, D, m, Hp, H, pc))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty(M, H, device=cuda)
    rc = lib.ctr_cin_layer_fwd(X0c.data_ptr(), ld0, Xpc.data_ptr(), Hp, Wc.data_ptr(), bc.data_ptr(),
                               B, D, m, Hp, H, out.data_ptr(), pc, ws.data_ptr(), ws.numel(), st)
    assert rc == 0, _lib.last_error()
    dX0 = torch.zeros(M, ld0, device=cuda)
    dXp = torch.ones(M, Hp, device=cuda)          # += semantics: starts at 1
    dW = torch.zeros(m * Hp, H, device=cuda)
    db = torch.zeros(H, device=cuda)
    rc = lib.ctr_cin_layer_bwd(X0c.data_ptr(), ld0, Xpc.data_ptr(), Hp, Wc.data_ptr(), dc.data_ptr(),
                               B, D, m, Hp, H, dX0.data_ptr(), dXp.data_ptr(), dW.data_ptr(),
                               db.data_ptr(), pc, ws.data_ptr(), ws.numel(), st)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    x0, xp, w, d = X0t[:, :m].double(), Xp.double(), W.double().view(m, Hp, H), dpre.double()
    pre = torch.einsum("ri,rj,ijh->rh", x0, xp, w) + bias.double()
    tol = CIN_TOL[prec]
    bad = []

    def close(a, b, what):
        err = float((a.cpu().double() - b).abs().max())
        s = float(b.abs().max()) + 1e-9
        if not err <= tol * s:
            bad.append("%s: max err %.3e vs scale %.3e" % (what, err, s))
    # compare where the pre-activation is not within tolerance of the ReLU kink
    ref = torch.relu(pre)
    safe = pre.abs() > 10 * tol * float(pre.abs().max())
    close(out.cpu().double() * safe, ref * safe, "out")
    close(dX0[:, :m], torch.einsum("rj,rh,ijh->ri", xp, d, w), "dX0t")
    close(dXp - 1.0, torch.einsum("ri,rh,ijh->rj", x0, d, w), "dXp")
    close(dW, torch.einsum("ri,rj,rh->ijh", x0, xp, d).reshape(m * Hp, H), "dW")
    close(db, d.sum(0), "dbias")
    assert not bad, "%s: %s" % (prec, "; ".join(bad))


@pytest.mark.parametrize("prec", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("B,m,D,layers", [(9, 39, 16, (128, 128)), (64, 39, 16, (20, 10, 10)),
                                          (300, 39, 16, (128, 128)), (17, 5, 8, (16,))])
def test_cin_fwd_bwd(cuda, prec, B, m, D, layers):
    """The whole CIN stack through autograd.  The pooled output is checked in max-norm; the
    gradients in the 99.5th percentile because an element whose pre-activation sits within
    rounding error of 0 can flip its ReLU mask (a discontinuity, not a kernel error)."""
    ops = _ops()
    g = torch.Generator().manual_seed(B + m)
    E = torch.randn(B, m * D, generator=g) * 0.25
    Ws, bs = [], []
    hp = m
    for h in layers:
        Ws.append(torch.randn(m * hp, h, generator=g) * (2.0 / (m * hp + h)) ** 0.5)
        bs.append(torch.randn(h, generator=g) * 0.01)
        hp = h
    dp = torch.randn(B, sum(layers), generator=g)
    Ec = E.to(cuda).requires_grad_(True)
    Wc = [w.to(cuda).requires_grad_(True) for w in Ws]
    bc = [b.to(cuda).requires_grad_(True) for b in bs]
    out = ops.cin(Ec, m, D, Wc, bc, prec)
    out.backward(dp.to(cuda))
    Eo = E.double().requires_grad_(True)
    Wo = [w.double().requires_grad_(True) for w in Ws]
    bo = [b.double().requires_grad_(True) for b in bs]
    ref = _cin_oracle(Eo, Wo, bo, m, D)
    ref.backward(dp.double())
    tol = {"fp32": 5e-5, "tf32x3": 5e-5, "tf32": 2e-2}[prec]
    bad = []

    def close(a, b, what, q=1.0):
        e = (a.detach().cpu().double() - b.detach()).abs().reshape(-1)
        err = float(e.max()) if q >= 1.0 else float(torch.quantile(e[:: max(1, e.numel() // 200000)], q))
        s = float(b.detach().abs().max()) + 1e-9
        if not err <= tol * s:
            bad.append("%s: err(q=%.3f) %.3e vs scale %.3e" % (what, q, err, s))
    close(out, ref, "pooled")
    close(Ec.grad, Eo.grad, "dE", q=0.995)
    for k in range(len(layers)):
        close(Wc[k].grad, Wo[k].grad, "dW%d" % k, q=0.995)
        close(bc[k].grad, bo[k].grad, "db%d" % k, q=0.9)
    assert not bad, "%s: %s" % (prec, "; ".join(bad))


def test_transpose_roundtrip(cuda):
    from recsys_b200 import _lib
    lib = _lib.load()
    B, F, D, ld = 37, 39, 16, 40
    E = torch.randn(B, F * D, device=cuda)
    Xt = torch.full((B * D, ld), 7.0, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.ctr_transpose_fd(E.data_ptr(), B, F, D, Xt.data_ptr(), ld, st) == 0
    want = E.view(B, F, D).permute(0, 2, 1).reshape(B * D, F)
    assert torch.equal(Xt[:, :F], want) and float(Xt[:, F:].abs().max()) == 0.0
    dE = torch.ones(B, F * D, device=cuda)
    assert lib.ctr_transpose_df_add(Xt.data_ptr(), ld, B, F, D, dE.data_ptr(), st) == 0
    assert torch.allclose(dE, E + 1.0)


# -------------------------------------------------------------------------- DIN
@pytest.mark.parametrize("B,P,E", [(1, 100, 16), (33, 100, 16), (257, 37, 16), (40, 100, 32),
                                   (40, 64, 8)])
def test_din_attention_fwd_bwd(cuda, B, P, E):
    ops = _ops()
    from recsys_b200 import feature_column as fc
    n_items = 500
    feats, _ = synth.synthetic_din(B, P=P, seed=B, n_items=n_items, n_cates=50)
    if B > 1:
        feats["u_iid_seq"][1] = 0                       # a sample whose whole history is padding
    cols = [fc.embedding_column(fc.categorical_column_with_hash_bucket("i_id", n_items), E)]
    lay = fc.Layout(cols, ["i_id"], [n_items], [0, n_items], E)
    emb = ops.FieldEmbedding(lay, cuda, with_w1=False, seed=2, record=False)   # din.cu: planar rows
    g = torch.Generator().manual_seed(B)
    p64 = {}
    sizes = [4 * E, 80, 40, 1]
    for l in range(3):
        p64[f"att.{l}.w"] = (torch.randn(sizes[l], sizes[l + 1], generator=g) *
                             (2.0 / (sizes[l] + sizes[l + 1])) ** 0.5).double()
        p64[f"att.{l}.b"] = (torch.randn(sizes[l + 1], generator=g) * 0.1).double()
    query = torch.randn(B, E, generator=g) * 0.5
    dout = torch.randn(B, E, generator=g)
    hist = torch.from_numpy(feats["u_iid_seq"])
    pc = {k: v.float().to(cuda).requires_grad_(True) for k, v in p64.items()}
    qc = query.to(cuda).requires_grad_(True)
    out = ops.din_attention(emb, 0, hist.to(cuda, torch.int32), qc, pc["att.0.w"], pc["att.0.b"],
                            pc["att.1.w"], pc["att.1.b"], pc["att.2.w"], pc["att.2.b"])
    out.backward(dout.to(cuda))
    po = {k: v.clone().requires_grad_(True) for k, v in p64.items()}
    tab = emb.table.double().cpu().requires_grad_(True)
    qo = query.double().requires_grad_(True)
    ref = om.din_attention(po, "att", tab, hist, qo, False, 0.0, None)
    ref.backward(dout.double())

    def close(a, b, what, tol=2e-4):
        err = float((a.detach().cpu().double() - b.detach()).abs().max())
        s = float(b.detach().abs().max()) + 1e-9
        assert err <= tol * s + 1e-6, "%s: max err %.3e vs scale %.3e" % (what, err, s)

    close(out, ref, "out")
    close(qc.grad, qo.grad, "dquery")
    close(emb.dtable, tab.grad, "dtable")
    for k in p64:
        close(pc[k].grad, po[k].grad.reshape(pc[k].grad.shape), "d" + k)
