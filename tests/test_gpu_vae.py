"""AutoencoderKL on the tcgen05 kernels (mvedit_b200.vae) vs the fp32 oracle (oracle/vae_oracle.py, diffusers architecture restated:
PARITY UNPINNED -- diffusers and the SD-1.5 VAE weights are absent offline; random-init weights of the published shapes).
Tolerance: ||out - ref||_2 / ||ref||_2 <= 3e-2 (bf16 storage through ~30 convolutions), the same bar as the UNet tests."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize('cfg_name,B,L', [('tiny', 3, 16), ('sd15', 2, 16), ('sd15', 1, 64)])
def test_decode_matches_fp32_oracle(cfg_name, B, L):
    from oracle import vae_oracle as vo
    from mvedit_b200.vae import AutoencoderKL, VAEConfig
    ocfg = vo.TINY_VAE if cfg_name == 'tiny' else vo.SD15_VAE
    sd = {k: v.cuda() for k, v in vo.random_vae_state_dict(ocfg, seed=3).items()}
    vae = AutoencoderKL(sd, VAEConfig(**ocfg.__dict__))
    g = torch.Generator(device='cuda').manual_seed(L + B)
    z = torch.randn(B, 4, L, L, device='cuda', generator=g)
    with torch.no_grad():
        out = vae.decode(z, return_dict=False)[0]
        ref = vo.decode(sd, ocfg, z)
        assert out.shape == ref.shape == (B, 3, 8 * L, 8 * L)
        assert _rel(out, ref) < 3e-2, _rel(out, ref)
        # the reference's decode tail (mvedit_3d_pipeline.py:1258-1263) in one call
        img = vae.decode_images(z * ocfg.scaling_factor)
        ref_img = vo.decode_targets(sd, ocfg, z * ocfg.scaling_factor)
        assert img.dtype == torch.float32 and img.shape == (B, 8 * L, 8 * L, 3) and float(img.min()) >= 0 and float(img.max()) <= 1
        assert (img - ref_img).abs().mean().item() < 1.5e-2


@pytest.mark.parametrize('cfg_name,B,S', [('tiny', 2, 64), ('sd15', 1, 128)])
def test_encode_matches_fp32_oracle(cfg_name, B, S):
    from oracle import vae_oracle as vo
    from mvedit_b200.vae import AutoencoderKL, VAEConfig
    ocfg = vo.TINY_VAE if cfg_name == 'tiny' else vo.SD15_VAE
    sd = {k: v.cuda() for k, v in vo.random_vae_state_dict(ocfg, seed=4).items()}
    vae = AutoencoderKL(sd, VAEConfig(**ocfg.__dict__))
    g = torch.Generator(device='cuda').manual_seed(S)
    x = torch.rand(B, 3, S, S, device='cuda', generator=g) * 2 - 1
    with torch.no_grad():
        dist = vae.encode(x).latent_dist
        mean, logvar = vo.encode_moments(sd, ocfg, x)
    assert dist.mean.shape == (B, 4, S // 8, S // 8)
    assert _rel(dist.mean, mean) < 3e-2, _rel(dist.mean, mean)
    assert _rel(dist.logvar, logvar) < 3e-2
    s = dist.sample()
    assert s.shape == mean.shape and torch.isfinite(s).all()


def test_softmax_rows_and_transposed_value_gemm():
    """the single-head d=512 attention path: score GEMM -> mve_softmax_rows_bf16 -> value GEMM with V^T from a weight-as-A GEMM."""
    from mvedit_b200 import tc_ops as T
    g = torch.Generator(device='cuda').manual_seed(0)
    S, C = 256, 128
    x = torch.randn(S, 1024, device='cuda', generator=g) * 3
    xb = x.to(torch.bfloat16)
    out = T.softmax_rows(xb.clone(), 0.37)
    ref = torch.softmax(xb.float() * 0.37, dim=-1)
    assert (out.float() - ref).abs().max().item() < 2e-3 and abs(out.float().sum(-1).mean().item() - 1) < 2e-2
    # strided rows (a column slice of a wider buffer)
    wide = torch.zeros(S, 2048, dtype=torch.bfloat16, device='cuda')
    wide[:, :1024] = xb
    out2 = T.softmax_rows(wide[:, :1024], 0.37, out=torch.empty(S, 1024, dtype=torch.bfloat16, device='cuda'))
    assert torch.equal(out2, out)
    h = torch.randn(2 * S, C, device='cuda', generator=g).to(torch.bfloat16)
    wv = (torch.randn(C, C, device='cuda', generator=g) / C ** 0.5).to(torch.bfloat16)
    vt = T.gemm(wv, h)                                            # [C, 2S] = (h @ wv^T)^T
    torch.testing.assert_close(vt.float(), (h.float() @ wv.float().t()).t(), rtol=2e-2, atol=2e-2)
    p = torch.softmax(torch.randn(S, S, device='cuda', generator=g), -1).to(torch.bfloat16)
    o = T.gemm(p, vt[:, S:2 * S])                                 # B operand = a column block of V^T (ldb = 2S)
    torch.testing.assert_close(o.float(), p.float() @ vt[:, S:2 * S].float().t(), rtol=2e-2, atol=2e-2)
