"""Stage-by-stage check of ONE decode step of a 1-layer small model against the oracle (debug aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import build_engine, engine_config, f32, make_weights, oracle_configs
from oracle import restate as R
import torch.nn.functional as F

dt = torch.bfloat16
P = int(sys.argv[1]) if len(sys.argv) > 1 else 37
cfg = engine_config(dt, vit_layers=0, proj_d_model=0, gate_layers=0, llm_layers=1, use_graphs=False)
sd = make_weights(cfg, vit=False, proj=False, gate=False, llm=True)
eng = build_engine(cfg, sd)
oc = oracle_configs(cfg).llm
sd32 = f32(sd)
g = torch.Generator().manual_seed(3)
ids = torch.randint(3, cfg.llm_vocab, (P,), generator=g)
emb = eng.embed_tokens(ids.cuda())
eng.llm_prefill(emb)
out = eng.llm_decode(2)
def buf(which, n):
    t = torch.empty(n, dtype=dt, device="cuda")
    eng._check(eng.lib.sm_debug_decode_buffer(eng._h, which, t.data_ptr(), n * 2, None))
    torch.cuda.synchronize()
    return t.float().cpu()
Hq, Hk, D, H, Fd = oc.num_heads, oc.num_kv_heads, oc.head_dim, oc.hidden_size, oc.intermediate_size
with R.emulate(dt):
    cache = R.KVCache()
    lg0 = R.mistral_forward(sd32, "", oc, sd32["model.embed_tokens.weight"][ids], cache)
    t0 = int(lg0.argmax())
    print("first token", out[0], t0)
    x = sd32["model.embed_tokens.weight"][t0][None]
    lp = "model.layers.0."
    a = R.rms_norm(x, sd32[lp + "input_layernorm.weight"], oc.rms_norm_eps)
    q = R._linear(a, sd32[lp + "self_attn.q_proj.weight"]); k = R._linear(a, sd32[lp + "self_attn.k_proj.weight"]); v = R._linear(a, sd32[lp + "self_attn.v_proj.weight"])
    qkv_o = torch.cat([q, k, v], -1)[0]
    cos, sin = R.rope_cos_sin(oc, torch.tensor([P]))
    qh = R.apply_rope(q.view(1, Hq, D).transpose(0, 1), cos, sin); kh = R.apply_rope(k.view(1, Hk, D).transpose(0, 1), cos, sin)
    kk = torch.cat([cache.k[0], kh], 1).repeat_interleave(Hq // Hk, 0); vv = torch.cat([cache.v[0], v.view(1, Hk, D).transpose(0, 1)], 1).repeat_interleave(Hq // Hk, 0)
    s = (qh @ kk.transpose(-1, -2)) * D ** -0.5
    pr = R._r(torch.softmax(s, -1))
    att_o = R._r((pr @ vv).transpose(0, 1).reshape(Hq * D))
    h1 = R._r(x[0] + R._linear(att_o, sd32[lp + "self_attn.o_proj.weight"]))
    a2 = R.rms_norm(h1, sd32[lp + "post_attention_layernorm.weight"], oc.rms_norm_eps)
    m_o = R._r(R._r(F.silu(R._linear(a2, sd32[lp + "mlp.gate_proj.weight"]))) * R._linear(a2, sd32[lp + "mlp.up_proj.weight"]))
    h2 = R._r(h1 + R._linear(m_o, sd32[lp + "mlp.down_proj.weight"]))
for name, got, exp in (("qkv", buf(1, (Hq + 2 * Hk) * D), qkv_o), ("att", buf(2, Hq * D), att_o), ("m", buf(3, Fd), m_o), ("x", buf(0, H), h2)):
    d = (got - exp).abs()
    print(f"{name}: max abs diff {d.max():.4g} at {int(d.argmax())}, ref max {exp.abs().max():.4g}, rel L2 {(d.norm() / exp.norm()):.3g}")
    if name == "att":
        for h in range(Hq):
            dh = d[h * D:(h + 1) * D]
            print(f"   head {h}: max diff {dh.max():.4g}")
eng.close()
