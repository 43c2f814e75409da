"""CPU-only tests: the C-ABI library loads and exports every declared symbol, parameter layout and
state_dict compatibility, optimizer state layout, the drop-in import paths, error behaviour."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import groove_oracle as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "groove_b200.h")).read()
    declared = set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        getattr(lib, name)
    assert lib.gt_version() == 1


def test_param_layout_matches_reference_state_dict_order():
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    for cfg in (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), G.GrooveCfg(24, 3, 40, 2, 3, 27, 27)):
        c = _lib.GtConfig(cfg.d_model, cfg.nhead, cfg.dim_ff, cfg.n_enc, cfg.n_dec, cfg.e_src, cfg.e_tgt, 0, 0.0, 0)
        shapes = G.param_shapes(cfg)
        offs, sizes = (C.c_int64 * len(shapes))(), (C.c_int64 * len(shapes))()
        assert lib.gt_param_layout(C.byref(c), offs, sizes, len(shapes)) == len(shapes)
        assert [int(s) for s in sizes] == [int(np.prod(s)) for _, s in shapes]
        assert all(int(o) % 4 == 0 for o in offs) and list(offs) == sorted(offs)
        assert lib.gt_param_count(C.byref(c)) >= sum(int(s) for s in sizes)


def test_invalid_configs_are_rejected_with_messages():
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    bad = _lib.GtConfig(30, 4, 16, 1, 0, 16, 27, 0, 0.0, 0)
    assert lib.gt_param_count(C.byref(bad)) < 0 and b"divisible" in lib.gt_last_error()
    bad = _lib.GtConfig(32, 4, 16, 1, 0, 16, 26, 0, 0.0, 0)
    assert lib.gt_workspace_bytes(C.byref(bad), 4, 1) < 0 and b"multiple of 3" in lib.gt_last_error()
    ok = _lib.GtConfig(32, 4, 16, 1, 0, 16, 27, 0, 0.0, 0)
    assert lib.gt_workspace_bytes(C.byref(ok), 0, 1) < 0
    assert lib.gt_workspace_bytes(C.byref(ok), 8, 1) > lib.gt_workspace_bytes(C.byref(ok), 8, 0) > 0
    assert lib.gt_sgd_step(None, None, 4, 0.1, 1.0, None) != 0        # null pointers are refused before any launch


def test_precision_modes_select_their_paths():
    """gt_path_kind is host logic: fp32 -> FFMA kernels, fp32_tc -> the per-op kernels with split-operand tcgen05 contractions for
    every shape (its workspace carries the three-image operand scratch), bf16 -> fused kernels where they exist; anything else is
    refused."""
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    cfg = lambda prec, d=32, h=16, f=512, ld=0: _lib.GtConfig(d, h, f, 6, ld, 16, 27, prec, 0.1, 0)
    assert lib.gt_path_kind(C.byref(cfg(_lib.PREC_FP32))) == _lib.PATH_FP32_SIMT
    assert lib.gt_path_kind(C.byref(cfg(_lib.PREC_BF16))) == _lib.PATH_FUSED_D32
    for shape in (dict(), dict(d=256, h=2), dict(d=256, h=16, f=64), dict(ld=6), dict(d=64, h=4, f=128)):
        assert lib.gt_path_kind(C.byref(cfg(_lib.PREC_FP32_TC, **shape))) == _lib.PATH_GEMM_TC_SPLIT
    w32, wtc = (lib.gt_workspace_bytes(C.byref(cfg(p, d=256, h=2)), 64, 1) for p in (_lib.PREC_FP32, _lib.PREC_FP32_TC))
    assert wtc > w32 > 0
    assert lib.gt_path_kind(C.byref(cfg(3))) < 0 and b"precision" in lib.gt_last_error()
    from transformergrooveinfilling_b200 import GrooveTransformerEncoder
    m = GrooveTransformerEncoder(32, 16, 27, 4, 64, 0.1, 2, 32, "cpu")
    assert m.set_precision("fp32_tc")._cfg().precision == _lib.PREC_FP32_TC
    with pytest.raises(ValueError):
        m.set_precision("tf32")


def test_state_dict_keys_match_reference_demo_checkpoint():
    from BaseGrooveTransformers.models.transformer import GrooveTransformerEncoder
    want = [l.split() for l in open(os.path.join(ROOT, "tests", "golden", "demo_checkpoint_keys.txt")) if not l.startswith("#")]
    m = GrooveTransformerEncoder(32, 16, 27, 4, 16, 0.18, 6, 32, "cpu")
    sd = m.state_dict()
    assert list(sd) == [w[0] for w in want]
    for w in want:
        assert list(sd[w[0]].shape) == [int(v) for v in w[1:]], w[0]
    np.testing.assert_allclose(sd["InputLayerEncoder.PositionalEncoding.pe"].numpy(), G.positional_table(32).numpy(), atol=1e-7)


def test_encdec_state_dict_and_flat_views():
    from BaseGrooveTransformers.models.transformer import GrooveTransformer
    cfg = G.GrooveCfg(24, 3, 40, 2, 3, 27, 27)
    m = GrooveTransformer(24, 27, 27, 3, 40, 0.1, 2, 3, 32, "cpu")
    names = [k for k, _ in G.param_shapes(cfg)]
    assert [k for k in m.state_dict() if not k.endswith(".pe")] == names
    assert [n for n, _ in m.named_parameters()] == names
    # parameters are views of ONE flat vector: writing through a parameter changes the flat vector
    with torch.no_grad():
        m.OutputLayer.Linear.bias.fill_(3.0)
    assert float(m.flat_parameters().detach()[-28:-1].sum()) == 81.0
    P = G.det_params(cfg)
    sd = m.state_dict(); sd.update(P); m.load_state_dict(sd, strict=True)
    o, s = m._offsets[5]
    assert torch.equal(m.flat_parameters()[o:o + s].detach().view(P[names[5]].shape), P[names[5]])
    # reference init: all layers of a stack start identical; in/out layers U(-0.1,0.1) with zero bias
    m2 = GrooveTransformer(32, 27, 27, 4, 64, 0.1, 3, 2, 32, "cpu")
    assert torch.equal(m2.Encoder.Encoder.layers[0].linear1.weight, m2.Encoder.Encoder.layers[2].linear1.weight)
    assert float(m2.OutputLayer.Linear.weight.abs().max()) <= 0.1 and float(m2.OutputLayer.Linear.bias.abs().max()) == 0
    assert float(m2.Decoder.Decoder.layers[1].norm3.weight.mean()) == 1.0


def test_dropin_imports_and_signatures():
    import inspect
    from BaseGrooveTransformers import calculate_loss, initialize_model, train_loop
    assert list(inspect.signature(train_loop).parameters) == [
        "dataloader", "groove_transformer", "loss_fn", "bce_fn", "mse_fn", "opt", "epoch", "save", "device",
        "encoder_only", "hit_loss_penalty", "test_inputs", "test_gt", "validation_inputs", "validation_gt"]
    assert list(inspect.signature(calculate_loss).parameters) == ["prediction", "y", "bce_fn", "mse_fn", "hit_loss_penalty"]
    params = {"model": dict(encoder_only=1, optimizer="sgd", d_model=32, n_heads=4, dim_feedforward=16, dropout=0.1,
                            num_encoder_layers=2, num_decoder_layers=0, max_len=32, embedding_size_src=16,
                            embedding_size_tgt=27, device="cpu"),
              "training": dict(learning_rate=0.05, batch_size=4), "load_model": None}
    model, opt, epoch = initialize_model(params)
    assert epoch == 0 and type(model).__name__ == "GrooveTransformerEncoder"
    sd = opt.state_dict()
    assert sd["param_groups"][0]["lr"] == 0.05 and sd["param_groups"][0]["params"] == list(range(len(list(model.parameters()))))
    # the SGD state-dict has the key set of the reference's demo checkpoint optimizer
    line = [l for l in open(os.path.join(ROOT, "tests", "golden", "demo_checkpoint_keys.txt")) if l.startswith("#")][0]
    for key in ("dampening", "lr", "momentum", "nesterov", "params", "weight_decay"):
        assert key in sd["param_groups"][0] and key in line
    params["model"]["optimizer"] = "adam"; params["model"]["encoder_only"] = 0; params["model"]["num_decoder_layers"] = 1
    model, opt, _ = initialize_model(params)
    assert type(model).__name__ == "GrooveTransformer" and type(opt).__name__ == "FusedAdam"
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 32, 16), torch.zeros(1, 32, 27))     # CPU tensors: the product has no CPU path


def test_checkpoint_resume_roundtrip(tmp_path):
    from BaseGrooveTransformers import initialize_model
    params = {"model": dict(encoder_only=1, optimizer="adam", d_model=16, n_heads=2, dim_feedforward=8, dropout=0.0,
                            num_encoder_layers=1, num_decoder_layers=0, max_len=32, embedding_size_src=16,
                            embedding_size_tgt=27, device="cpu"),
              "training": dict(learning_rate=0.01, batch_size=4), "load_model": None}
    model, opt, _ = initialize_model(params)
    opt._t = 3; opt._m.fill_(0.5)
    torch.save({"epoch": 7, "model_state_dict": model.state_dict(), "optimizer_state_dict": opt.state_dict(), "loss": 1.0},
               tmp_path / "transformer_run_abc_Epoch_7.Model")
    params["load_model"] = {"location": "local", "dir": str(tmp_path), "file_pattern": "transformer_run_{}_Epoch_{}.Model"}
    m2, o2, ep = initialize_model(params)
    assert ep == 7 and o2._t == 3 and float(o2._m[:64].mean()) == 0.5
    assert torch.equal(m2.flat_parameters().detach(), model.flat_parameters().detach())
    params["load_model"]["dir"] = str(tmp_path / "nothing")
    with pytest.raises(FileNotFoundError):
        initialize_model(params)


def test_grad_buckets_partition_the_flat_vector_in_backward_order():
    """gt_grad_buckets: contiguous ranges in completion order (head first, encoder input layer last) that tile the
    flat gradient exactly; layer buckets start at that layer's first tensor."""
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    for cfg in (G.GrooveCfg(32, 16, 512, 6, 0, 16, 27), G.GrooveCfg(24, 3, 40, 2, 3, 27, 27)):
        c = _lib.GtConfig(cfg.d_model, cfg.nhead, cfg.dim_ff, cfg.n_enc, cfg.n_dec, cfg.e_src, cfg.e_tgt, 0, 0.0, 0)
        total = lib.gt_param_count(C.byref(c))
        offs, sizes = (C.c_int64 * 160)(), (C.c_int64 * 160)()
        n = lib.gt_grad_buckets(C.byref(c), offs, sizes, 160)
        assert n == (cfg.n_enc + cfg.n_dec + 3 if cfg.n_dec else cfg.n_enc + 2)
        end = total
        for i in range(n):
            assert offs[i] + sizes[i] == end and sizes[i] > 0
            end = offs[i]
        assert end == 0
        names = [k for k, _ in G.param_shapes(cfg)]
        po, ps = (C.c_int64 * len(names))(), (C.c_int64 * len(names))()
        lib.gt_param_layout(C.byref(c), po, ps, len(names))
        start = {names[i]: int(po[i]) for i in range(len(names))}
        first_enc_bucket = cfg.n_dec + 2 if cfg.n_dec else 1
        assert offs[first_enc_bucket] == start[f"Encoder.Encoder.layers.{cfg.n_enc - 1}.self_attn.in_proj_weight"]
        assert offs[n - 1] == 0 and offs[n - 2] == start["Encoder.Encoder.layers.0.self_attn.in_proj_weight"]
        assert lib.gt_grad_buckets(C.byref(c), offs, sizes, 1) < 0
    assert lib.gt_grad_bucket_wait(0, None) != 0 and b"not enabled" in lib.gt_last_error()


def test_graph_entry_points_validate_their_arguments_before_any_launch():
    """gt_graph_train_create / gt_graph_launch / gt_graph_destroy (include/groove_b200.h): argument errors come back as a
    status + message, never as a launch — checked here without a GPU."""
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    assert lib.gt_graph_launch(None, 1, None) != 0 and b"gt_graph_launch" in lib.gt_last_error()
    assert lib.gt_graph_destroy(None) == 0
    handle = C.c_void_p()
    fp32 = _lib.GtConfig(32, 4, 16, 1, 0, 16, 27, _lib.PREC_FP32, 0.1, 0)
    # null buffers are refused first; every path is capturable, so what stops the other calls is the workspace check of their own path
    assert lib.gt_graph_train_create(C.byref(fp32), None, None, None, None, 4, 0.5, None, None, None, None, 0, 0, 0.1, None, None, 1,
                                     None, None, None, None, None, 0, None, C.byref(handle)) != 0
    assert b"null pointer" in lib.gt_last_error() and not handle.value
    one = C.c_void_p(16)        # any non-null address: validation stops at the (misaligned) workspace, nothing is dereferenced
    d256 = _lib.GtConfig(256, 16, 64, 1, 0, 16, 27, _lib.PREC_BF16, 0.1, 0)        # fused d_model = 256 path
    encdec = _lib.GtConfig(32, 4, 16, 1, 1, 16, 27, _lib.PREC_BF16, 0.1, 0)        # encoder-decoder on the fused d_model = 32 path
    for cfg in (d256, encdec, fp32):
        assert lib.gt_graph_train_create(C.byref(cfg), one, one, one, one, 4, 0.5, one, one, one, one, 1 << 20, 0, 0.1, None, None, 1,
                                         one, None, None, None, None, 0, None, C.byref(handle)) != 0
        assert b"256-byte aligned" in lib.gt_last_error() and not handle.value
    # Adam without its moment vectors is refused by name
    assert lib.gt_graph_train_create(C.byref(d256), one, one, one, one, 4, 0.5, one, one, one, one, 1 << 20, 1, 0.1, None, None, 1,
                                     one, None, None, None, None, 0, None, C.byref(handle)) != 0
    assert b"optimizer must be" in lib.gt_last_error() and not handle.value


def test_input_pipelines_refuse_a_cpu_device():
    from transformergrooveinfilling_b200.pipeline import DeviceResidentLoader, HostBatchPrefetcher
    with pytest.raises(RuntimeError):
        HostBatchPrefetcher("cpu", (4, 32, 16), (4, 32, 27))
    with pytest.raises(RuntimeError):
        DeviceResidentLoader(torch.zeros(4, 32, 16), torch.zeros(4, 32, 27), 2, "cpu")


def test_sample_sweep_and_params_from_config():
    """sweep.py host logic: draws follow the wandb spec of configs/InfillingClosedHH_sweep.yaml and map to train.py's params."""
    from transformergrooveinfilling_b200.sweep import params_from_config, sample_sweep
    spec = {"method": "random", "parameters": {
        "batch_size": {"values": [16, 32, 64]}, "d_model": {"values": [16, 32, 64, 128]}, "dim_feedforward": {"values": [16, 512]},
        "dropout": {"distribution": "uniform", "min": 0.1, "max": 0.3}, "optimizer_algorithm": {"value": "sgd"},
        "learning_rate": {"distribution": "uniform", "min": 0, "max": 0.1}, "n_heads": {"values": [1, 2, 4, 8, 16, 32]},
        "num_encoder_decoder_layers": {"distribution": "int_uniform", "min": 6, "max": 12}, "epochs": {"value": 100},
        "encoder_only": {"value": 1}, "experiment": {"value": "InfillingClosedHH"},
        "hit_loss_penalty": {"distribution": "uniform", "min": 0, "max": 1}}}
    a, b = sample_sweep(spec, 20, seed=3), sample_sweep(spec, 20, seed=3)
    assert a == b and len(a) == 20
    for c in a:
        assert c["d_model"] % c["n_heads"] == 0 and 6 <= c["num_encoder_decoder_layers"] <= 12
        assert 0.1 <= c["dropout"] <= 0.3 and c["optimizer_algorithm"] == "sgd"
        p = params_from_config(c, "cuda", "bf16")
        assert p["model"]["num_encoder_layers"] == c["num_encoder_decoder_layers"] and p["model"]["num_decoder_layers"] == 0
        assert p["model"]["embedding_size_src"] == 16 and p["model"]["embedding_size_tgt"] == 27 and p["model"]["max_len"] == 32
        assert p["training"]["batch_size"] == c["batch_size"] and p["load_model"] is None
    sym = dict(a[0], experiment="InfillingClosedHH_Symbolic", encoder_only=0)
    p = params_from_config(sym)
    assert p["model"]["embedding_size_src"] == 27 and p["model"]["num_decoder_layers"] == sym["num_encoder_decoder_layers"]


def test_bench_flop_accounting_matches_survey():
    """bench.py's algorithmic FLOPs per sequence are the SURVEY.md §8(d) figures (2 FLOP per MAC, backward = 2x forward)."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    saved = (os.dup(1), sys.stdout)                     # bench.py points fd 1 at stderr on import; undo that for pytest
    try:
        spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved[0], 1)
        sys.stdout = saved[1]
    want = {"c1": 8.522e6, "c2": 45.091e6, "c3": 624.968e6, "c4": 659.571e6, "c5": 97.229e6}
    for k, v in want.items():
        assert abs(bench.train_flops_per_seq(bench.WORKLOADS[k]) - v) / v < 1e-3, k
    for k, w in bench.WORKLOADS.items():                # hyper-parameters verbatim from the reference yamls
        assert w["d"] % w["H"] == 0 and w["batch"] % 4 == 0
    x, y = bench.synth_batch(bench.WORKLOADS["c2"], 8, 1)
    assert x.shape == (8, 32, 16) and y.shape == (8, 32, 27)
    hits = y[..., :9]
    assert set(hits.unique().tolist()) <= {0.0, 1.0}
    assert float((y[..., 9:18] * (1 - hits)).abs().max()) == 0.0 and float(y[..., 18:].abs().max()) <= 0.5


def test_deepcopy_and_pickle_keep_parameters_bound_to_the_flat_vector():
    """copy.deepcopy (best-model snapshots, EMA) and pickling must give a model whose nn.Parameters are still views of the ONE
    flat vector the kernels read (the default deepcopy detached them silently)."""
    import copy
    import pickle
    from transformergrooveinfilling_b200 import GrooveTransformerEncoder
    m = GrooveTransformerEncoder(32, 16, 27, 4, 16, 0.18, 2, 32, "cpu")
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert torch.equal(clone.flat_parameters().detach(), m.flat_parameters().detach())
        assert clone.flat_parameters().data_ptr() != m.flat_parameters().data_ptr()
        with torch.no_grad():
            clone.OutputLayer.Linear.bias.fill_(2.0)                 # write through a parameter ...
        assert float(clone.flat_parameters().detach()[-28:-1].sum()) == 54.0     # ... lands in the clone's flat vector
        assert float(m.flat_parameters().detach()[-28:-1].abs().sum()) == 0.0      # ... and not in the original's
        sd = m.state_dict()
        sd["Encoder.Encoder.norm.weight"] = torch.full((32,), 0.5)
        clone.load_state_dict(sd)
        o, s = clone._offsets[[n for n, _ in clone._spec_list].index("Encoder.Encoder.norm.weight")]
        assert torch.equal(clone.flat_parameters().detach()[o:o + s], torch.full((32,), 0.5))
        assert [n for n, _ in clone.named_parameters()] == [n for n, _ in m.named_parameters()]


def test_dropout_seed_follows_torch_seed():
    from transformergrooveinfilling_b200 import GrooveTransformerEncoder
    torch.manual_seed(123)
    a = GrooveTransformerEncoder(32, 16, 27, 4, 16, 0.18, 1, 32, "cpu")
    torch.manual_seed(124)
    b = GrooveTransformerEncoder(32, 16, 27, 4, 16, 0.18, 1, 32, "cpu")
    assert a._seed == 123 and b._seed == 124 and a._step == 0


def test_calculate_loss_rejects_other_loss_functions():
    from transformergrooveinfilling_b200 import calculate_loss
    z = torch.zeros(2, 32, 9)
    with pytest.raises(ValueError, match="BCEWithLogitsLoss"):
        calculate_loss((z, z, z), torch.zeros(2, 32, 27), torch.nn.BCEWithLogitsLoss(), torch.nn.MSELoss(reduction="none"), 1.0)
    with pytest.raises(ValueError, match="MSELoss"):
        calculate_loss((z, z, z), torch.zeros(2, 32, 27), None, torch.nn.L1Loss(reduction="none"), 1.0)


def test_data_parallel_exchange_argument_is_validated():
    """exchange = 'nccl' | 'p2p' | 'auto': unknown values and p2p with an injected compute are refused before anything is set up; a
    single process (world 1) never builds a peer exchange."""
    from transformergrooveinfilling_b200.dp import DataParallelStep
    from transformergrooveinfilling_b200 import FusedSGD, GrooveTransformerEncoder
    m = GrooveTransformerEncoder(32, 16, 27, 4, 64, 0.1, 1, 32, "cpu")
    opt = FusedSGD(m, 0.1)
    with pytest.raises(ValueError, match="exchange"):
        DataParallelStep(m, opt, 0.5, exchange="rdma")
    with pytest.raises(ValueError, match="injected compute"):
        DataParallelStep(m, opt, 0.5, compute=lambda x, y: (None, None), exchange="p2p")
    dp = DataParallelStep(m, opt, 0.5, exchange="auto")
    assert dp.world == 1 and dp.peer is None and dp.exchange == "nccl" and hasattr(opt, "step_peers")
    # the C entry points refuse null / oversized arguments before any launch
    from transformergrooveinfilling_b200 import _lib
    lib = _lib.load()
    bufs = (C.c_void_p * 2)(None, None)
    assert lib.gt_sgd_step_peers(None, bufs, 2, 0, None, 4, 0.1, 1.0, None) != 0
    assert lib.gt_adam_step_peers(C.c_void_p(256), bufs, 2, 0, C.c_void_p(256), C.c_void_p(256), None, 4, 0.1, 0.9, 0.999, 1e-8, 1, 1.0, None) != 0
    assert b"exchange buffer" in lib.gt_last_error()
    assert lib.gt_adam_step_peers(C.c_void_p(256), bufs, 17, 0, C.c_void_p(256), C.c_void_p(256), None, 4, 0.1, 0.9, 0.999, 1e-8, 1, 1.0, None) != 0
    assert b"world size" in lib.gt_last_error()
    assert lib.gt_peer_publish(None, 0, None, 4, None) != 0 and lib.gt_peer_open(None, None) != 0
