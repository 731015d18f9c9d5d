"""GPU parity of the aggregation path (C ABI through the drop-in functions) vs the golden vectors generated
from the reference and vs the oracle.  Tolerance: 1e-5 relative on descriptors (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import segvlad_oracle as O
from revisit_anything_b200 import engine, func_vpr, synth
from revisit_anything_b200._lib import TOKENS_DN, TOKENS_ND

pytestmark = pytest.mark.gpu
RTOL = 1e-5   # north_star: VLAD descriptors within 1e-5 relative
AGG_CASES = ["agg_small_o2", "agg_small_o0", "agg_small_S3", "agg_unnorm_o1", "agg_realvocab_o3"]


def _cmp(got, want):
    """Elementwise 1e-5 relative; elements that are ~0 by cancellation are held to an absolute floor of
    1e-6 x the row's largest element instead (their relative error is unbounded for ANY fp32 pipeline:
    the token normalise x/||x|| alone differs by 1 ulp between two fp32 implementations)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    floor = 1e-6 * np.abs(want).max(axis=-1, keepdims=True)
    err = np.abs(got - want)
    bad = err > RTOL * np.abs(want) + floor
    assert not bad.any(), f"{int(bad.sum())} elements out of tolerance, max abs err {err.max():.3e}"


@pytest.mark.parametrize("name", AGG_CASES)
def test_golden_reference_vectors(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = {"desired_height": int(g["H"]), "desired_width": int(g["W"])}
    adj = torch.from_numpy(g["adj"]) if int(g["order"]) else None
    D = g["centers"].shape[1]
    gd = func_vpr.seg_vlad_gpu_single_img(None, None, torch.from_numpy(g["tokens"]), "img", list(g["masks"]),
                                          torch.from_numpy(g["centers"]), cfg, desc_dim=D, adj_mat=adj)
    assert gd.dtype == torch.float64 and not gd.is_cuda and gd.shape[0] == len(g["masks"])
    _cmp(gd.numpy()[:, g["cols"]], g["vlad_cols"])
    _cmp(np.linalg.norm(gd.numpy().reshape(gd.shape[0], 32, D), axis=2), g["row_block_norms"])


def _image(seed, D, H, W, S, K=32, order=3):
    dh, dw = H // 14, W // 14
    centers = synth.make_centers(K, D, seed)
    tokens = synth.make_tokens(D, dh, dw, seed, centers)
    masks = synth.make_masks(S, H // 2, W // 2, seed)
    adj = torch.from_numpy(O.neighbour_adjacency(masks, order)) if order else None
    return centers, tokens, masks, adj


def _check_against_oracle(seed, D, H, W, S, K, order, layout=TOKENS_DN, out_dtype=torch.float64):
    centers, tokens, masks, adj = _image(seed, D, H, W, S, K, order)
    cfg = {"desired_height": H, "desired_width": W}
    want, labels, margin, member = O.seg_vlad_single_img(tokens, masks, centers, cfg, adj)
    dev = torch.device("cuda")
    N = (H // 14) * (W // 14)
    bits = engine.mask_to_membership(torch.from_numpy(np.asarray(masks)).to(dev), H, W)
    np.testing.assert_array_equal(engine.pack_membership(member.to(dev)).cpu().numpy(), bits.cpu().numpy())
    tok = tokens.reshape(D, N).to(dev)
    if layout == TOKENS_ND:
        tok = tok.t().contiguous()
    got, lab = engine.aggregate_batch(tok, N, D, layout, centers.to(dev), bits, [S], [adj] if adj is not None else None,
                                      out_dtype=out_dtype, return_labels=True)
    safe = margin.numpy() > 1e-5
    np.testing.assert_array_equal(lab.cpu().numpy()[0][safe], labels.numpy()[safe])
    assert safe.mean() > 0.99
    if out_dtype == torch.float32:
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-7)
    else:
        _cmp(got.cpu().numpy(), want.numpy())


def test_17places_shape_vs_oracle():
    # 640x480 -> N = 1530 tokens, D_t = 1536, K = 32, order-3 SuperSegments (config 1 shape)
    _check_against_oracle(seed=101, D=1536, H=480, W=640, S=37, K=32, order=3)


@pytest.mark.parametrize("D,K,order,S", [(768, 32, 3, 21), (768, 64, 1, 9), (256, 128, 2, 17), (64, 32, 0, 5)])
def test_variants_vs_oracle(D, K, order, S):
    _check_against_oracle(seed=7 + D + K, D=D, H=196, W=266, S=S, K=K, order=order)


def test_north_star_table_k64_d1536_vs_oracle():
    # BASELINE north_star: 64 x 1536 centroid table at the 640x480 token grid
    _check_against_oracle(seed=164, D=1536, H=480, W=640, S=24, K=64, order=3)


def test_config5_table_k128_d768_vs_oracle():
    _check_against_oracle(seed=1128, D=768, H=322, W=322, S=19, K=128, order=2)


def test_canonical_150_segments_multi_tile_vs_oracle():
    # S = 150 (the survey's canonical image): two segment tiles (128 + 22), the 32-row store boxes end inside the image
    _check_against_oracle(seed=150, D=512, H=480, W=640, S=150, K=32, order=3)


def test_batch_of_tiles_at_image_boundaries_vs_oracle():
    # [150, 129, 128, 1]: tile tails of 22 / 1 / 0 rows and a single-segment image share one launch; a bulk store box
    # must never cross into the next image's rows
    dev = torch.device("cuda")
    D, H, W, K = 128, 224, 308, 32
    N = (H // 14) * (W // 14)
    cfg = {"desired_height": H, "desired_width": W}
    centers = synth.make_centers(K, D, 8)
    toks, bits, counts, adjs, wants = [], [], [], [], []
    for i, S in enumerate([150, 129, 128, 1]):
        tokens = synth.make_tokens(D, H // 14, W // 14, 400 + i, centers)
        masks = synth.make_masks(S, H // 2, W // 2, 500 + i)
        adj = torch.from_numpy(O.neighbour_adjacency(masks, 2))
        want, _, margin, _ = O.seg_vlad_single_img(tokens, masks, centers, cfg, adj)
        assert float(margin.min()) > 1e-5
        toks.append(tokens.reshape(D, N)); counts.append(S); adjs.append(adj); wants.append(want.numpy())
        bits.append(engine.mask_to_membership(torch.from_numpy(np.asarray(masks)).to(dev), H, W))
    for dt in (torch.float64, torch.float32):
        got = engine.aggregate_batch(torch.stack(toks).to(dev), N, D, TOKENS_DN, centers.to(dev), torch.cat(bits), counts,
                                     adjs, out_dtype=dt).cpu().numpy()
        _cmp(got, np.concatenate(wants))


@pytest.mark.parametrize("frac", [0.6, 0.95])
@pytest.mark.parametrize("D,K", [(1536, 32), (256, 64)])
def test_skewed_vocabulary_one_dominant_cluster(frac, D, K):
    # one cluster holds 60 % / 95 % of the image's 1530 tokens (~900 / ~1450 tokens): the accumulator chain of such a
    # cluster is split into bounded sub-chains (csrc/aggregate_tc.cuh kTcSubChunks), the error must not grow with n_k
    dev = torch.device("cuda")
    H, W, S = 480, 640, 12
    N = (H // 14) * (W // 14)
    cfg = {"desired_height": H, "desired_width": W}
    centers = synth.make_centers(K, D, 77)
    tokens = synth.make_tokens_skewed(D, H // 14, W // 14, 31 + K, centers, frac, heavy=3)
    masks = synth.make_masks(S, H // 2, W // 2, 33)
    adj = torch.from_numpy(O.neighbour_adjacency(masks, 3))
    want, labels, margin, _ = O.seg_vlad_single_img(tokens, masks, centers, cfg, adj)
    assert int((labels == 3).sum()) > 0.9 * frac * N
    safe = margin.numpy() > 1e-5
    assert safe.all()
    bits = engine.mask_to_membership(torch.from_numpy(np.asarray(masks)).to(dev), H, W)
    got, lab = engine.aggregate_batch(tokens.reshape(D, N).to(dev), N, D, TOKENS_DN, centers.to(dev), bits, [S], [adj],
                                      return_labels=True)
    np.testing.assert_array_equal(lab.cpu().numpy()[0], labels.numpy())
    _cmp(got.cpu().numpy(), want.numpy())
    # the AnyLoc whole-image path feeds ALL tokens of the image to one segment (utilities.VLAD, S = 1)
    from revisit_anything_b200.utilities import VLAD
    v = VLAD(K, desc_dim=None, dist_mode="cosine", vlad_mode="hard", cache_dir=None)
    v.set_centers(centers)
    tok_nd = tokens.reshape(D, N).t().contiguous()
    _cmp(v.generate(tok_nd).numpy()[None], O.anyloc_vlad_generate(tok_nd, centers).numpy()[None])


def test_seg_vlad_gpu_single_with_mapping():
    # the function north_star names (func_vpr.py:1065-1101): tokens come out of an h5py-like mapping
    H, W, D, K, S = 196, 266, 96, 32, 7
    cfg = {"desired_height": H, "desired_width": W}
    centers, tokens, masks, adj = _image(5, D, H, W, S, K, 2)
    store = {"img_000.jpg": {"ift_dino": tokens.numpy()}}
    gd = func_vpr.seg_vlad_gpu_single(None, None, store, "img_000.jpg", masks, centers, cfg, desc_dim=D, adj_mat=adj)
    assert gd.dtype == torch.float64 and not gd.is_cuda and tuple(gd.shape) == (S, K * D)
    want = O.seg_vlad_single_img(tokens, masks, centers, cfg, adj)[0]
    _cmp(gd.numpy(), want.numpy())
    twin = func_vpr.seg_vlad_gpu_single_img(None, None, tokens, "img_000.jpg", masks, centers, cfg, desc_dim=D, adj_mat=adj)
    np.testing.assert_array_equal(gd.numpy(), twin.numpy())


def test_token_major_layout_and_f32_output():
    _check_against_oracle(seed=55, D=384, H=196, W=266, S=12, K=32, order=2, layout=TOKENS_ND)
    _check_against_oracle(seed=56, D=384, H=196, W=266, S=12, K=32, order=2, out_dtype=torch.float32)


def test_batched_images_ragged_segments_and_edge_cases():
    # images with S = 1, 2, 3 (no Delaunay), an EMPTY mask row (zero vector), and a normal one, in one launch
    dev = torch.device("cuda")
    D, H, W, K = 128, 140, 182, 32
    N = (H // 14) * (W // 14)
    cfg = {"desired_height": H, "desired_width": W}
    centers = synth.make_centers(K, D, 3)
    toks, bits, counts, adjs, wants = [], [], [], [], []
    for i, S in enumerate([1, 2, 3, 0, 11, 8]):
        tokens = synth.make_tokens(D, H // 14, W // 14, 200 + i, centers)
        masks = synth.make_masks(S, H // 2, W // 2, 300 + i) if S else []
        if S == 8:
            masks[2][:] = False                      # empty segment (cannot come out of SAM, must still work)
            adj = None
        else:
            adj = torch.from_numpy(O.neighbour_adjacency(masks, 2)) if S else None
        toks.append(tokens.reshape(D, N))
        counts.append(S)
        adjs.append(adj)
        if S:
            want, _, margin, member = O.seg_vlad_single_img(tokens, masks, centers, cfg, adj)
            assert float(margin.min()) > 1e-5
            wants.append(want.numpy())
            bits.append(engine.mask_to_membership(torch.from_numpy(np.asarray(masks)).to(dev), H, W))
    got = engine.aggregate_batch(torch.stack(toks).to(dev), N, D, TOKENS_DN, centers.to(dev), torch.cat(bits), counts,
                                 adjs).cpu().numpy()
    want = np.concatenate(wants)
    _cmp(got, want)
    empty_row = 1 + 2 + 3 + 11 + 2
    assert np.abs(got[empty_row]).max() == 0.0


def test_zero_residual_blocks_use_exact_row_norm():
    # tokens exactly equal to (unit-norm) centres -> residual rows are exactly 0 -> block norm 0 although the
    # cluster is populated: exercises the rownorm_fixup kernel (reference: F.normalize of an all-zero block)
    dev = torch.device("cuda")
    D, K, N, S = 64, 32, 96, 4
    g = torch.Generator().manual_seed(0)
    centers = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1)
    x = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    x[:20] = centers[3]
    x[20:30] = centers[7]
    member = torch.rand(S, N, generator=g) < 0.4
    member[0, :30] = True
    member[0, 30:] = False                           # segment 0 sees only zero-residual tokens -> all-zero row
    lab0, _ = O.assign_labels(x, centers)
    member[1] = (lab0 != 3) & (torch.rand(N, generator=g) < 0.5)
    member[1, :20] = True                            # block (1, 3) is populated, but only by zero residuals
    assert int(((lab0 == 3) & member[1]).sum()) == 20
    want, labels, margin = O.vlad_single(x, centers, member, None)
    out, _ = func_vpr.vlad_single(x.to(dev), centers.to(dev), None, member.to(dev), None)
    _cmp(out.cpu().numpy(), want.numpy())
    assert np.abs(out[0].cpu().numpy()).max() == 0.0


def test_vlad_matmuls_per_cluster_dropin():
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(1)
    N, D, K, S = 150, 96, 32, 10
    res = torch.randn(N, D, generator=g)
    labels = torch.randint(0, K, (N,), generator=g)
    masks = (torch.rand(S, N, generator=g) < 0.2)
    adj = (torch.rand(S, S, generator=g) < 0.3) | torch.eye(S, dtype=torch.bool)
    want = O.aggregate_per_cluster(K, masks, res, labels, adj)
    got, _ = func_vpr.vlad_matmuls_per_cluster(K, masks.double().to(dev), res.double().to(dev), labels.to(dev),
                                              adjMat=adj.double().to(dev))
    _cmp(got.cpu().numpy(), want.numpy())
    with pytest.raises(RuntimeError):
        func_vpr.vlad_matmuls_per_cluster(K, masks.double(), res.double(), labels, device="cpu")


@pytest.mark.parametrize("H,W,Hm,Wm", [(480, 640, 240, 320), (480, 640, 480, 640), (224, 300, 100, 133)])
def test_mask_to_membership_kernel(H, W, Hm, Wm):
    masks = synth.make_masks(9, Hm, Wm, 5)
    masks[0][:] = False
    masks[0][-1, -1] = True                          # single pixel in the folded-in border cell
    want = O.mask_to_patch_membership(masks, H, W)
    dev = torch.device("cuda")
    bits = engine.mask_to_membership(torch.from_numpy(np.asarray(masks)).to(dev), H, W)
    np.testing.assert_array_equal(bits.cpu().numpy(),
                                  engine.pack_membership(torch.from_numpy(want).to(dev)).cpu().numpy())


def test_mask_centroids_feed_the_same_adjacency():
    # GPU centroids (integer sums / count in fp64) are the doubles of np.nonzero(m).mean (func_vpr.py:1314); the adjacency
    # built from them equals the host-only drop-in and the oracle; an empty mask gives NaN like numpy
    dev = torch.device("cuda")
    for S, Hm, Wm, order in [(37, 240, 320, 3), (9, 100, 133, 2), (3, 48, 64, 1)]:
        masks = synth.make_masks(S, Hm, Wm, 60 + S)
        c = engine.mask_centroids(torch.from_numpy(np.stack(masks)).to(dev)).cpu().numpy()
        want = np.array([np.array(np.nonzero(m)).mean(1)[::-1] for m in masks])
        np.testing.assert_array_equal(c, want)
        a_gpu = func_vpr.nbrMasksAGGFastSingle(masks, order, centroids=c)
        assert torch.equal(a_gpu, func_vpr.nbrMasksAGGFastSingle(masks, order))
        np.testing.assert_array_equal(a_gpu.numpy(), O.neighbour_adjacency(masks, order))
    empty = torch.zeros((2, 8, 8), dtype=torch.uint8, device=dev)
    empty[1, 3, 5] = 1
    c = engine.mask_centroids(empty).cpu().numpy()
    assert np.isnan(c[0]).all() and c[1].tolist() == [5.0, 3.0]


def test_bad_arguments_raise():
    dev = torch.device("cuda")
    with pytest.raises(ValueError):
        engine.aggregate_batch(torch.zeros(30 * 6, device=dev), 30, 6, TOKENS_DN, torch.zeros(32, 6, device=dev),
                               torch.zeros((1, 1), dtype=torch.int32, device=dev), [1])   # D % 4 != 0


@pytest.mark.parametrize("B,D,K,H,W", [(3, 1536, 32, 480, 640), (2, 768, 128, 322, 322), (2, 1520, 64, 196, 266),
                                       (1, 64, 8, 140, 182)])
def test_tensor_core_assignment_matches_simt_assignment(monkeypatch, B, D, K, H, W):
    # labels / ||x|| from the tcgen05 kernel (csrc/assign_tc.cu: three bf16 pieces, six products) against the SIMT fp32
    # kernels (SEGVLAD_ASSIGN_TC=0) and the oracle's fp64 argmax; D = 1520 ends in a partial 64-channel stage, N = 529 is odd
    dev = torch.device("cuda")
    dh, dw = H // 14, W // 14
    N = dh * dw
    centers = synth.make_centers(K, D, 40 + K)
    toks = torch.stack([synth.make_tokens(D, dh, dw, 900 + i, centers).reshape(D, N) for i in range(B)])
    if D == 64:
        toks = toks * 37.5                                   # un-normalised tokens: ||x|| is computed by the kernel
    S = 5
    member = torch.rand(B * S, N, generator=torch.Generator().manual_seed(1)) < 0.5
    bits = engine.pack_membership(member.to(dev))
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SEGVLAD_ASSIGN_TC", mode)
        outs[mode] = engine.aggregate_batch(toks.to(dev), N, D, TOKENS_DN, centers.to(dev), bits, [S] * B, None,
                                            return_labels=True)
    lab_tc, lab_simt = outs["1"][1].cpu().numpy(), outs["0"][1].cpu().numpy()
    n_safe = 0
    for b in range(B):
        want, margin = O.assign_labels(torch.nn.functional.normalize(toks[b].t().contiguous(), dim=1), centers)
        safe = margin.numpy() > 1e-5
        n_safe += int(safe.sum())
        np.testing.assert_array_equal(lab_tc[b][safe], want.numpy()[safe])
        np.testing.assert_array_equal(lab_simt[b][safe], want.numpy()[safe])
    assert n_safe > 0.99 * B * N
    if (lab_tc == lab_simt).all():
        _cmp(outs["1"][0].cpu().numpy(), outs["0"][0].cpu().numpy())
