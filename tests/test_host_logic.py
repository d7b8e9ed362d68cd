"""CPU: host-side mirror of the reference interface (names, signatures, state_dict layout, errors)."""
import inspect
import os
from argparse import Namespace

import numpy as np
import pytest
import torch

import swnerf_b200 as S
from swnerf_b200 import dnerf
from oracle import nerf_oracle as O


def test_state_dict_layout_matches_reference_shapes():
    m = S.vallina_NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    sd = m.state_dict()
    ref = O.mlp_param_shapes()
    assert list(sd.keys()) == O.mlp_param_names() or set(sd.keys()) == set(ref.keys())
    for k, sh in ref.items():
        assert tuple(sd[k].shape) == sh, k
    assert sum(p.numel() for p in m.parameters()) == 595844
    d = S.DirectTemporalNeRF(D=8, W=256, input_ch=63, input_ch_views=27, input_ch_time=21, output_ch=5,
                             skips=[4], use_viewdirs=True, embed_fn=None)
    refd = O.dnerf_param_shapes()
    sdd = d.state_dict()
    assert set(sdd.keys()) == set(refd.keys())
    for k, sh in refd.items():
        assert tuple(sdd[k].shape) == sh, k
    assert sum(p.numel() for p in d.parameters()) == 1095047
    nv = S.vallina_NeRF(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=4, skips=[4], use_viewdirs=False)
    assert set(nv.state_dict().keys()) == set(O.mlp_param_shapes(input_ch_views=0, use_viewdirs=False).keys())


def test_param_list_order_matches_header():
    m = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True)
    names = {id(p): n for n, p in m.named_parameters()}
    got = [names[id(p)] for p in m.param_list()]
    want = [f"pts_linears.{i}.{w}" for i in range(8) for w in ("weight", "bias")] + \
           ["views_linears.0.weight", "views_linears.0.bias", "feature_linear.weight", "feature_linear.bias",
            "alpha_linear.weight", "alpha_linear.bias", "rgb_linear.weight", "rgb_linear.bias"]
    assert got == want


def test_same_default_init_as_reference_constructor_order():
    # nn.Linear containers created in the reference's order -> identical RNG consumption
    torch.manual_seed(0)
    a = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(63, 256)
    assert torch.equal(a.pts_linears[0].weight, lin.weight)


def test_get_embedder_dims():
    for L, d, want in [(10, 3, 63), (4, 3, 27), (10, 1, 21), (20, 3, 123), (8, 1, 17), (4, 1, 9)]:
        fn, od = S.get_embedder(L, d, 0)
        assert od == want == O.embed_dim(L, d)
        assert fn.L == L and fn.input_dims == d
    fn, od = S.get_embedder(-1, 3, -1)
    assert od == 3 and isinstance(fn, torch.nn.Identity)
    x = torch.randn(4, 3)
    assert torch.equal(fn(x), x)


def test_signatures_match_reference():
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(S.render_rays) == ["ray_batch", "network_fn", "network_query_fn", "N_samples", "retraw", "lindisp",
                                  "perturb", "N_importance", "network_fine", "white_bkgd", "raw_noise_std",
                                  "verbose", "pytest"]
    assert sig(dnerf.render_rays) == sig(S.render_rays) + ["z_vals", "use_two_models_for_fine"]
    assert sig(S.run_network) == ["inputs", "viewdirs", "fn", "embed_fn", "embeddirs_fn", "netchunk"]
    assert sig(dnerf.run_network) == ["inputs", "viewdirs", "frame_time", "fn", "embed_fn", "embeddirs_fn",
                                      "embedtime_fn", "netchunk", "embd_time_discr"]
    assert sig(S.sample_pdf) == ["bins", "weights", "N_samples", "det", "pytest"]
    assert sig(S.raw2outputs) == ["raw", "z_vals", "rays_d", "raw_noise_std", "white_bkgd", "pytest"]
    assert sig(S.get_embedder) == ["multires", "input_dims", "i"]


def _args(tmp, **kw):
    a = dict(multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=128, N_samples=64,
             netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, lrate=5e-4,
             ft_path=None, basedir=str(tmp), expname="e", no_reload=False, perturb=1.0, white_bkgd=True,
             raw_noise_std=0.0, dataset_type="blender", no_ndc=False, lindisp=False, nerf_type="direct_temporal",
             use_two_models_for_fine=False, not_zero_canonical=False, do_half_precision=False)
    a.update(kw)
    os.makedirs(os.path.join(str(tmp), "e"), exist_ok=True)
    return Namespace(**a)


def test_create_nerf_contract_and_checkpoint_roundtrip(tmp_path):
    args = _args(tmp_path)
    kw_train, kw_test, start, grad_vars, opt = S.create_nerf(args, device=torch.device("cpu"))
    assert start == 0 and len(grad_vars) == 48
    assert sum(p.numel() for p in grad_vars) == 1191688
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    assert kw_train["ndc"] is False and set(kw_train) >= {"network_query_fn", "network_fn", "network_fine",
                                                          "N_samples", "N_importance", "white_bkgd"}
    assert isinstance(opt, torch.optim.Adam)
    # checkpoint in the reference's format (nerf/run.py:716-724) reloads
    torch.save({"global_step": 7, "network_fn_state_dict": kw_train["network_fn"].state_dict(),
                "network_fine_state_dict": kw_train["network_fine"].state_dict(),
                "optimizer_state_dict": opt.state_dict()}, os.path.join(str(tmp_path), "e", "000007.tar"))
    _, _, start2, _, _ = S.create_nerf(args, device=torch.device("cpu"))
    assert start2 == 7


def test_create_nerf_dnerf_and_multires(tmp_path):
    args = _args(tmp_path)
    kw, _, _, gv, _ = dnerf.create_nerf(args, device=torch.device("cpu"))
    assert sum(p.numel() for p in gv) == 1095047 and kw["network_fine"] is None
    for ch, dims in [((20, 8, 20), (123, 17, 123)), ((10, 4, 10), (63, 9, 63)), ((-1, -1, -1), (3, 1, 3))]:
        kw, _, _, _, _ = dnerf.create_nerf_multires(args, ch, 0, device=torch.device("cpu"))
        m = kw["network_fn"]
        assert (m.input_ch, m.input_ch_time, m.input_ch_views) == dims     # SURVEY 3.4 [probe]


def test_no_cpu_path():
    m = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(4, 90))
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.raw2outputs(torch.zeros(2, 8, 4), torch.zeros(2, 8), torch.ones(2, 3))
    with pytest.raises(AssertionError):
        S.searchsorted(torch.zeros(2, 3), torch.zeros(3, 3))             # searchsorted.py:25-27 row rule


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sw-nerf_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_tnerf_module_layout_and_signatures():
    """model.TNeRF / tnerf.* keep the reference's names, shapes and signatures (model.py:152-210, run_tnerf.py)."""
    from swnerf_b200 import tnerf
    m = S.TNeRF(depth=8, in_feat=63, dir_feat=27, time_feat=21, net_dim=128, skip_layer=4)
    ref = O.tnerf_param_shapes()
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())                       # same registration order as the reference
    for k, sh in ref.items():
        assert tuple(sd[k].shape) == sh, k
    assert tuple(m.layers[5][0].weight.shape) == (128, 128 + 63 + 21)   # the re-injection after layer 4
    assert isinstance(m.layers[0][1], torch.nn.ELU) and isinstance(m.color[1], torch.nn.ReLU)
    assert m.spec.act == "elu" and m.spec.skip_extra and m.spec.rgb_relu and m.spec.skips == (4,)
    assert [tuple(p.shape) for p in m.param_list()[-8:]] == [(64, 155), (64,), (128, 128), (128,), (1, 128), (1,),
                                                            (3, 64), (3,)]
    torch.manual_seed(0)
    a = S.TNeRF(8, 63, 27, 21)
    torch.manual_seed(0)
    lin = torch.nn.Linear(84, 128)
    assert torch.equal(a.layers[0][0].weight, lin.weight)            # same RNG consumption order
    with pytest.raises(ValueError):
        S.TNeRF(depth=10, in_feat=63, dir_feat=27, time_feat=21)     # inconsistent in the reference itself
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(tnerf.render_rays) == sig(dnerf.render_rays)
    assert sig(tnerf.run_network) == ["inputs", "viewdirs", "frame_time", "fn", "embed_fn", "embeddirs_fn",
                                      "embedtime_fn", "netchunk", "embd_time_discr"]
    assert sig(m.forward) == ["inp", "vdir", "dyn_t"]
    with pytest.raises((RuntimeError, TypeError)):                   # no CPU path
        m(torch.zeros(4, 90), torch.zeros(4, 27), torch.zeros(4, 21))


def test_precision_switch_arms_layerwise_tc_gemm():
    """'tc' precision routes shapes without a fused kernel to the layer-wise tcgen05 GEMMs (spec.tc), 'fp32' to the fp32
    SIMT check path; the fused-kernel shape is untouched by the flag."""
    from swnerf_b200 import tnerf
    e3, e1 = S.get_embedder(10, 3, 0)[0], S.get_embedder(10, 1, 0)[0]
    ev = S.get_embedder(4, 3, 0)[0]
    m = S.TNeRF(8, 63, 27, 21)
    assert m.spec.tc is False
    for prec, want in (("tc", True), ("fp32", False)):
        tnerf.TNerfNetworkQuery(e3, ev, e1, precision=prec)._arm(m)
        assert m.tc_gemm is want and m.spec.tc is want
        v = S.vallina_NeRF(8, 256, 63, 0, 4, [4], False)
        S.NetworkQuery(e3, None, precision=prec)._arm(v)
        assert v.spec.tc is want
        d = S.DirectTemporalNeRF(D=8, W=256, input_ch=63, input_ch_views=63, input_ch_time=9, output_ch=5, skips=[4],
                                 use_viewdirs=True, embed_fn=e3)
        dnerf.DNerfNetworkQuery(e3, S.get_embedder(10, 3, 0)[0], S.get_embedder(4, 1, 0)[0], precision=prec)._arm(d)
        assert d.tc_gemm is want and d.time_spec.tc is want
    with pytest.raises(ValueError):
        tnerf.TNerfNetworkQuery(e3, ev, e1, precision="bf16")


def test_synthetic_workload_matches_the_oracles_generators():
    """bench.py and tools/ take their inputs from swnerf_b200.synth (the product never imports the oracle); the cpu_baseline
    leg takes them from the oracle: both must be the same rays and the same scene, bit for bit."""
    from swnerf_b200 import synth
    assert np.array_equal(synth.blender_rays(257, 7), O.blender_rays(257, 7))
    assert np.array_equal(synth.blender_rays(33, 3, frame_time=0.4), O.blender_rays(33, 3, frame_time=0.4))
    assert np.array_equal(synth.pose_spherical(40.0, -30.0, 4.0), O.pose_spherical(40.0, -30.0, 4.0))
    m = S.vallina_NeRF(8, 256, 63, 27, 5, [4], True)
    p, q = synth.scene_params(m, 21), O.make_params(O.mlp_param_shapes(), 21)
    assert list(p) == list(q) and all(torch.equal(p[k], q[k]) for k in q)
    t = S.TNeRF(8, 63, 27, 21)
    p, q = synth.scene_params(t, 5), O.make_params(O.tnerf_param_shapes(), 5)
    assert all(torch.equal(p[k], q[k]) for k in q)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    body = src.split("def main():")[1]
    assert "oracle" not in body.split("cpu_reference_run(")[0].replace("oracle port", "")   # main arm: no oracle import
