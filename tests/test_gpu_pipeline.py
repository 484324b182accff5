"""GPU parity of the host-fed end-to-end API (engine.HostFedPipeline / step_from_host): every step uploads NEW
node features and a NEW base subspace; the result must follow the CPU oracle step by step, with eager launches
and under CUDA-graph replay (the captured step reads its inputs at fixed addresses)."""
import numpy as np
import pytest
import torch

from gpu_util import pkg, dev
from oracle import step_port

pytestmark = pytest.mark.gpu


def _small_problem(k=16, hidden=(64, 64), freq=12):
    fem, syn = pkg("fem"), pkg("synthetic")
    verts, tris = syn.icosphere(freq)
    verts = fem.normalize_verts(verts)
    K, M = fem.assemble_stiffness_mass(verts, tris)
    n = verts.shape[0]
    rng = np.random.default_rng(7)
    Y, _ = syn.real_spherical_harmonics(verts / np.linalg.norm(verts, axis=1)[:, None], k)
    ei = torch.from_numpy(fem.connectivity_edges(tris))
    xs, Us = [], []
    for step in range(5):                              # different inputs for every step
        xs.append(torch.from_numpy(rng.standard_normal((n, 9 + k)).astype(np.float32)))
        U0 = torch.from_numpy((Y + 0.05 * rng.standard_normal(Y.shape)).astype(np.float32))
        Us.append(step_port.m_normalize(U0, M))
    return dict(K=K, M=M, n=n, k=k, hidden=list(hidden), ei=ei, xs=xs, Us=Us)


def _oracle_losses(pb):
    tr = step_port.CorrectorTrainer(pb["xs"][0], pb["ei"], pb["Us"][0], [pb["K"]], [pb["M"]], torch.zeros(pb["k"]),
                                    pb["hidden"], pb["k"])
    tr.epoch = 2500
    w0 = [w.detach().clone() for w in tr.weights], [b.detach().clone() for b in tr.biases]
    out = []
    for x, U in zip(pb["xs"], pb["Us"]):
        tr.x, tr.U_base = x, U
        out.append(tr.step()[:3])
    return np.array(out), w0


def _engine(pb, w0, mlp_mode):
    ops, sparse, engine = pkg("ops"), pkg("sparse"), pkg("engine")
    adj = sparse.CsrMatrix.from_edge_index(pb["ei"], pb["n"], dev())
    h = ops.neighbor_mean_concat(pb["xs"][0].to(dev()), adj)
    params = engine.FlatParams(w0[0], w0[1], dev())
    eng = engine.TrainStepEngine(h, pb["Us"][0].to(dev()), [sparse.OperatorPair(pb["K"], pb["M"], dev())], [0], params,
                                 engine.StepConfig(), lam_target=torch.zeros(pb["k"]).to(dev()), mlp_mode=mlp_mode)
    return eng, adj


@pytest.mark.parametrize("mlp_mode,graph", [("fp32", False), ("fp32", True), ("bf16", False), ("bf16", True)])
def test_host_fed_pipeline_follows_oracle_with_changing_inputs(mlp_mode, graph):
    ops, engine = pkg("ops"), pkg("engine")
    pb = _small_problem()
    want, w0 = _oracle_losses(pb)
    eng, adj = _engine(pb, w0, mlp_mode)
    pipe = engine.HostFedPipeline(eng, lambda x, out: ops.neighbor_mean_concat(x, adj, out=out))
    xs = [x.pin_memory() for x in pb["xs"]]
    Us = [U.pin_memory() for U in pb["Us"]]
    got = []
    u_ptr = eng.U_base.data_ptr()
    for i, (x, U) in enumerate(zip(xs, Us)):
        if graph and i == 1:
            eng.enable_graph()                           # steps 1.. are captured / replayed
        handle = pipe.submit(x, U, 2500 + i)
        got.append(pipe.result(handle)[[5, 0, 1]])
    got = np.array(got)
    assert eng.U_base.data_ptr() == u_ptr               # the engine's buffer is filled, never rebound
    tol = 1e-4 if mlp_mode == "fp32" else 2.5e-2
    np.testing.assert_allclose(got, want, rtol=tol)
    if mlp_mode == "bf16":
        np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=2e-3)
    # the five inputs really differ: the losses are not all the same
    assert np.ptp(want[:, 0]) > 1e-3 * want[:, 0].mean()


def test_step_from_host_equals_pipeline():
    ops, engine = pkg("ops"), pkg("engine")
    pb = _small_problem()
    want, w0 = _oracle_losses(pb)
    eng, adj = _engine(pb, w0, "fp32")
    got = []
    for i, (x, U) in enumerate(zip(pb["xs"], pb["Us"])):
        h = step_port.neighbour_mean_concat(x, pb["ei"]).pin_memory()
        got.append(eng.step_from_host(h, U.pin_memory(), 2500 + i)[[5, 0, 1]])
    np.testing.assert_allclose(np.array(got), want, rtol=1e-4)


def test_rebinding_inputs_of_a_captured_step_is_an_error():
    engine, cabi = pkg("engine"), pkg("_cabi")
    pb = _small_problem()
    _, w0 = _oracle_losses(pb)
    eng, _ = _engine(pb, w0, "fp32")
    eng.step(2500)
    eng.enable_graph()
    eng.step(2501)
    eng.U_base = eng.U_base.clone()
    with pytest.raises(cabi.EpError):
        eng.step(2502)
