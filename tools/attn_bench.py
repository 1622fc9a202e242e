import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipegoose_b200.ops import native, kernels as K
from pipegoose_b200.ops.attention import _AlibiAttentionNative
n = native()
for (B, S, H, D) in [(8, 1024, 16, 64), (4, 2048, 32, 128)]:
    qkv = torch.randn(B * S, H * 3 * D, device="cuda", dtype=torch.bfloat16)
    slopes = K.alibi_slopes(H, device="cuda")
    out = torch.empty(B * S, H * D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device="cuda")
    dout = torch.randn_like(out); dqkv = torch.empty_like(qkv)
    t0 = time.time(); n.attention_fwd(qkv, slopes, out, lse, B, S, H, D); torch.cuda.synchronize(); print("first fwd wall", time.time() - t0, flush=True)
    t0 = time.time(); n.attention_bwd(qkv, slopes, out, lse, dout, dqkv, B, S, H, D); torch.cuda.synchronize(); print("first bwd wall", time.time() - t0, flush=True)
    for name, fn in (("fwd", lambda: n.attention_fwd(qkv, slopes, out, lse, B, S, H, D)), ("bwd", lambda: n.attention_bwd(qkv, slopes, out, lse, dout, dqkv, B, S, H, D))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        fl = 4 * B * H * S * S * D / 2 * (1 if name == "fwd" else 2.5)
        print(B, S, H, D, name, f"{ms:.3f} ms", f"{fl / ms / 1e9:.1f} TFLOP/s", flush=True)
