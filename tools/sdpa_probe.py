"""Which SDPA backend is fastest for the two attention shapes of the ViT-B/16 dual encoder (harness question)."""
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

def bench(B, H, T, D, backend):
    q, k, v = (torch.randn((B, T, H, D), device="cuda", dtype=torch.bfloat16).transpose(1, 2).requires_grad_(True) for _ in range(3))
    do = torch.randn((B, H, T, D), device="cuda", dtype=torch.bfloat16)
    def run():
        with sdpa_kernel(backend):
            o = F.scaled_dot_product_attention(q, k, v)
        o.backward(do)
    try:
        for _ in range(3):
            run()
    except Exception as e:
        return f"unavailable ({type(e).__name__})"
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        run()
    b.record(); torch.cuda.synchronize()
    return f"{1e3 * a.elapsed_time(b) / 20:7.1f} us fwd+bwd"

for name, shape in [("vision 256x12x197x64", (256, 12, 197, 64)), ("text 256x8x77x64", (256, 8, 77, 64))]:
    for be in (SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION):
        print(f"{name:24s} {str(be):40s} {bench(*shape, be)}")


def bench_flash_attn(B, H, T, D):
    """flash_attn 2.x package (library code): takes (B, T, H, D) directly, the layout the towers' projections produce."""
    try:
        from flash_attn import flash_attn_func
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__})"
    q, k, v = (torch.randn((B, T, H, D), device="cuda", dtype=torch.bfloat16).requires_grad_(True) for _ in range(3))
    do = torch.randn((B, T, H, D), device="cuda", dtype=torch.bfloat16)
    def run():
        flash_attn_func(q, k, v).backward(do)
    try:
        for _ in range(3):
            run()
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__}: {str(e)[:80]})"
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        run()
    b.record(); torch.cuda.synchronize()
    return f"{1e3 * a.elapsed_time(b) / 20:7.1f} us fwd+bwd"


for name, shape in [("vision 256x12x197x64", (256, 12, 197, 64)), ("text 256x8x77x64", (256, 8, 77, 64)),
                    ("vit-l 256x16x257x64", (256, 16, 257, 64)), ("vit-l text 256x12x77x64", (256, 12, 77, 64))]:
    print(f"{name:24s} {'flash_attn package (flash_attn_func)':40s} {bench_flash_attn(*shape)}")
