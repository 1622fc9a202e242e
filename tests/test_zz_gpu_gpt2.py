"""GPT-2 on the fused bf16 kernels (GPU) against the same model in fp32 on the CPU.  Written without GPU access — the
kernels it uses are the validated Bloom ones (attention with zero ALiBi slopes, the gather kernel with apply_ln=False),
only their composition is new.  Collected last on purpose."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-6)).item()


def test_gpt2_matches_fp32_reference():
    from pipegoose_b200.models.gpt2 import GPT2Config, GPT2LMHeadModel

    torch.manual_seed(6)
    cfg = GPT2Config(vocab_size=4096, hidden_size=256, n_layer=2, n_head=4, n_positions=256)
    ref = GPT2LMHeadModel(cfg)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if "layernorm" in n or "ln_f" in n or "bias" in n:
                p.add_(torch.randn_like(p) * 0.1)
    model = copy.deepcopy(ref).to(torch.bfloat16).cuda()
    ref.load_state_dict({k: v.float().cpu() for k, v in model.state_dict().items()})
    ids = torch.randint(0, cfg.vocab_size, (2, 256))
    lr = ref(ids, labels=ids).loss
    lr.backward()
    lg = model(ids.cuda(), labels=ids.cuda()).loss
    lg.backward()
    assert abs(lg.item() - lr.item()) < 3e-2
    assert _rel(model(ids.cuda()).logits.cpu(), ref(ids).logits) < 3e-2
    rp = dict(ref.named_parameters())
    bad = [(n, _rel(p.grad.cpu(), rp[n].grad)) for n, p in model.named_parameters() if _rel(p.grad.cpu(), rp[n].grad) > 8e-2]
    assert not bad, bad
