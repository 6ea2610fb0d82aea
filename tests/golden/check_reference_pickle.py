# Build-container-only check (needs /root/reference; not a pytest): pickle the REAL reference network as main.py:1534-1536 does
# (`python check_reference_pickle.py save`), then unpickle it with this repo's model package (`... load`): same class path, flat
# parameter bit-identical to the reference's tensors.  Last run: all three checks True.
import sys, subprocess, types, torch, os
mode = sys.argv[1]
if mode == "save":
    sys.path.insert(0, "/root/reference")
    import model.nerf_raybased as ref
    torch.autograd.set_detect_anomaly(False)
    trial = types.SimpleNamespace(ON=True, body_arch="resmlp", res_scale=1.0, n_learnable=2, inact="relu", outact="none", n_block=-1, near=-1, far=-1)
    args = types.SimpleNamespace(netdepth=88, netwidth=256, layerwise_netwidths="", act="relu", linear_tail=False, use_residual=True, trial=trial)
    torch.manual_seed(0)
    m = ref.NeRF_v3_2(args, 1008, 3)
    torch.save({"network_fn": m, "network_fn_state_dict": m.state_dict()}, "/tmp/ref_ckpt.tar")
    print("saved", type(m).__module__)
else:
    sys.path.insert(0, "/root/repo")
    import model.nerf_raybased as ours
    ck = torch.load("/tmp/ref_ckpt.tar", weights_only=False)
    m = ck["network_fn"]
    print(type(m), m.flat.shape, list(m._modules), m.input_dim)
    from r2l_b200.nerf_raybased import init_flat_params
    print("flat == seed-0 reference weights:", torch.equal(m.flat.detach(), init_flat_params(0)))
    sd = m.state_dict()
    print("state_dict equal:", all(torch.equal(sd[k], v) for k, v in ck["network_fn_state_dict"].items()), len(sd))
    # round trip of our own pickles
    import io
    buf = io.BytesIO(); torch.save(m, buf); buf.seek(0); m2 = torch.load(buf, weights_only=False)
    print("own pickle round trip:", torch.equal(m2.flat, m.flat))
