"""Weight packing for the sm_100a kernels: diffusers-layout fp32/fp16 tensors -> K-contiguous GEMM operands
(fp16 for the tensor-core path; `dtype=torch.float32` for the fp32 validation path: same layouts)."""
import torch


def pack_conv3x3(w, dtype=torch.float16):
    """[Cout, Cin, 3, 3] -> [Cout, 9 * Cin_pad] ordered (ky, kx, cin); Cin zero-padded to a multiple of 64
    (only conv_in, Cin=4, needs padding: its activation operand is the 8-channel padded latent and TMA zero-fills
    the rest of the 64-wide K block)."""
    cout, cin = w.shape[0], w.shape[1]
    cin_pad = (cin + 63) // 64 * 64
    p = torch.zeros((cout, 3, 3, cin_pad), dtype=dtype, device=w.device)
    p[..., :cin] = w.permute(0, 2, 3, 1).to(dtype)
    return p.reshape(cout, 9 * cin_pad).contiguous()


def pack_linear(w, dtype=torch.float16):
    """[N, K] (nn.Linear) or [N, K, 1, 1] (1x1 conv) -> [N, K]."""
    return w.reshape(w.shape[0], -1).to(dtype).contiguous()


def pack_geglu(w, b, bn, return_perm=False, dtype=torch.float16):
    """GEGLU projection (ff.net.0.proj: [2F, K], chunk -> hidden | gate): interleave per BN-wide output tile as
    [BN/2 hidden rows | BN/2 matching gate rows] so the GEMM epilogue sees h_j and g_j of the same row in one
    TMEM accumulator tile. Returns (w_packed fp16 [2F, K], bias_packed fp32 [2F]) and, with `return_perm`, the row
    permutation (packed row i = original row perm[i]; a LoRA factor B is permuted the same way)."""
    two_f = w.shape[0]
    f = two_f // 2
    half = bn // 2
    if f % half != 0:
        raise ValueError(f"GEGLU width {f} not a multiple of BN/2={half}")
    idx = []
    for t in range(f // half):
        idx.extend(range(t * half, (t + 1) * half))
        idx.extend(range(f + t * half, f + (t + 1) * half))
    idx = torch.tensor(idx, device=w.device)
    wp, bp = w[idx].to(dtype).contiguous(), b[idx].to(torch.float32).contiguous()
    return (wp, bp, idx) if return_perm else (wp, bp)
