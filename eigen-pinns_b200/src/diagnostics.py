"""Post-training report: how close are the learned eigenpairs to the exact ones?

Drop-in for the metric part of reference src/diagnostics.py (:12-115 alignment helpers, :117-257 report): same
function names, arguments and return values, same printed tables.  Differences: the operators stay sparse (the
reference densifies K and M, `diagnostics.py:144-145`), the report is also RETURNED as a dict so that callers and tests
can use the numbers, and the 2 x 2 figure is drawn only when matplotlib is installed (it is not a dependency).
"""
import numpy as np
from scipy.linalg import svd
from scipy.optimize import linear_sum_assignment

import utils


def _overlap(A, B, M):
    return A.T @ (M @ B) if M is not None else A.T @ B


def align_eigenvectors(U_pred, U_exact, M=None, verbose=False):
    """Mode-by-mode matching: the assignment (Hungarian algorithm on |U_pred^T M U_exact|) that pairs every exact
    mode with one predicted mode, and the sign that makes each pair positively correlated.
    Returns (U_aligned, permutation, signs) with U_aligned[:, i] = signs[i] * U_pred[:, permutation[i]]."""
    k = U_pred.shape[1]
    W = np.asarray(_overlap(U_pred, U_exact, M))
    if verbose:
        print("\n=== Alignment Debug ===")
        print(f"Overlap shape: {W.shape}")
        print(f"Max abs overlap per row: {np.max(np.abs(W), axis=1)[:10]}")
    pred_idx, exact_idx = linear_sum_assignment(-np.abs(W))
    permutation = np.zeros(k, dtype=int)
    permutation[exact_idx] = pred_idx
    signs = np.sign(W[permutation, np.arange(k)])
    signs[signs == 0] = 1.0
    if verbose:
        for i in range(min(10, k)):
            print(f"  Exact {i} <- Predicted {permutation[i]} (overlap: {abs(W[permutation[i], i]):.4f})")
        print(f"\nSign flips: {signs[:10]}")
    return U_pred[:, permutation] * signs, permutation, signs


def get_subspace_error_and_alignment(U_pred, U_exact, M=None):
    """Orthogonal Procrustes: the rotation R = V D^T (from the SVD of U_pred^T M U_exact) minimising
    ||U_pred R - U_exact||_F.  Returns (U_pred @ R, that Frobenius norm)."""
    V, _, Dt = svd(np.asarray(_overlap(U_pred, U_exact, M)))
    U_aligned = U_pred @ (V @ Dt)
    return U_aligned, float(np.linalg.norm(U_aligned - U_exact, 'fro'))


def compute_rayleigh_quotients(U, L, M):
    """lambda_i = u_i^T L u_i / (u_i^T M u_i + 1e-12) for every column."""
    LU, MU = L @ U, M @ U
    return np.einsum("ij,ij->j", U, np.asarray(LU)) / (np.einsum("ij,ij->j", U, np.asarray(MU)) + 1e-12)


def comprehensive_diagnostics(U_pred, mesh, sampler, config):
    print("\n" + "=" * 80)
    print("COMPREHENSIVE EIGENMODE DIAGNOSTICS")
    print("=" * 80)
    print("Computing exact solution for comparison...")
    L, M = sampler.K_list[-1].tocsr(), sampler.M_list[-1].tocsr()
    if config.sampler_type == 'graph_coarsening':
        lambda_exact, U_exact, _, _ = utils.solve_eigenvalue_mesh(mesh, config.n_modes)
    elif config.sampler_type in ['farthest_point', 'voxel_downsampling']:
        lambda_exact, U_exact = utils.solve_eigenvalue_operators(L, M, config.n_modes)
    else:
        raise ValueError(f"Provided sampler type {config.sampler_type} is not supported!")
    print(f"Exact eigenvalues (first 10): {np.round(lambda_exact[:10], 6)}")
    lambda_pred = compute_rayleigh_quotients(U_pred, L, M)
    lambda_exact = compute_rayleigh_quotients(U_exact, L, M)
    print("\n--- BEFORE ALIGNMENT ---")
    print(f"Predicted eigenvalues (first 10): {np.round(lambda_pred[:10], 4)}")
    print(f"Exact eigenvalues (first 10): {np.round(lambda_exact[:10], 4)}")
    U_al, permutation, signs = align_eigenvectors(U_pred, U_exact, M, verbose=True)
    lambda_pred_ordered = lambda_pred[permutation]
    _, subspace_error = get_subspace_error_and_alignment(U_pred, U_exact, M)

    k = len(lambda_exact)
    shown = min(20, k)
    abs_err = np.abs(lambda_pred_ordered - lambda_exact)
    rel_err = abs_err / (np.abs(lambda_exact) + 1e-12)
    print("\n" + "=" * 80 + "\n1. EIGENVALUE COMPARISON\n" + "-" * 80)
    print(f"{'Mode':<6} {'λ_exact':<12} {'λ_pred':<12} {'Abs Err':<12} {'Rel Err':<12}\n" + "-" * 80)
    for i in range(shown):
        print(f"{i:<6} {lambda_exact[i]:<12.6f} {lambda_pred_ordered[i]:<12.6f} {abs_err[i]:<12.6f} {rel_err[i]:<12.6f}")
    print(f"\nSummary: Mean Abs Error = {abs_err.mean():.6f}, Mean Rel Error = {rel_err.mean():.6f}")

    MUa, MUe = M @ U_al, M @ U_exact
    inner = np.abs(np.einsum("ij,ij->j", U_al, MUe))
    cos = inner / (np.sqrt(np.einsum("ij,ij->j", U_al, MUa)) * np.sqrt(np.einsum("ij,ij->j", U_exact, MUe)) + 1e-12)
    l2 = np.linalg.norm(U_al - U_exact, axis=0)
    print("\n" + "=" * 80 + "\n2. EIGENVECTOR ALIGNMENT\n" + "-" * 80)
    print(f"{'Mode':<6} {'Perm':<10} {'Sign':<6} {'Cos Sim':<12} {'L2 Err':<12}\n" + "-" * 80)
    for i in range(shown):
        print(f"{i:<6} {i}->{permutation[i]:<7} {signs[i]:>4.0f}   {cos[i]:<12.6f} {l2[i]:<12.6f}")
    print(f"\nSummary: Mean Cos Sim = {cos.mean():.6f}, Mean L2 Error = {l2.mean():.6f}")
    print("\n" + "=" * 80 + "\n3. SUBSPACE ALIGNMENT (Procrustes)\n" + "-" * 80)
    print(f"Subspace Error (Frobenius): {subspace_error:.6e}")
    G = np.asarray(U_pred.T @ (M @ U_pred))
    diag = np.diag(G)
    off = float(np.max(np.abs(G - np.diag(diag))))
    print("\n" + "=" * 80 + "\n4. ORTHONORMALITY CHECK\n" + "-" * 80)
    print(f"Diagonal (first 10, should be ~1.0): {diag[:10]}")
    print(f"Max off-diagonal: {off:.6e}")
    if getattr(config, "diagnostics_viz", None):
        _create_diagnostic_plots(lambda_exact, lambda_pred_ordered, rel_err, cos, G, config)
    print("\n" + "=" * 80)
    return {"lambda_exact": lambda_exact, "lambda_pred": lambda_pred_ordered, "rel_errors": rel_err,
            "cos_sims": cos, "l2_errors": l2, "subspace_error": subspace_error, "permutation": permutation,
            "signs": signs, "gram": G, "max_off_diagonal": off}


def _create_diagnostic_plots(lambda_exact, lambda_pred, rel_errors, cos_sims, UMU, config):
    """2 x 2 overview (spectrum, relative errors, cosine similarities, Gram matrix); skipped without matplotlib."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        print("(matplotlib is not installed: diagnostic figure skipped)")
        return
    fig, ax = plt.subplots(2, 2, figsize=(12, 9))
    idx = np.arange(len(lambda_exact))
    ax[0, 0].plot(idx, lambda_exact, "o-", label="exact")
    ax[0, 0].plot(idx, lambda_pred, "x--", label="predicted")
    ax[0, 0].set_title("eigenvalues")
    ax[0, 0].legend()
    ax[0, 1].semilogy(idx, rel_errors + 1e-16, "o-")
    ax[0, 1].set_title("relative eigenvalue error")
    ax[1, 0].plot(idx, cos_sims, "o-")
    ax[1, 0].set_title("M-cosine similarity")
    im = ax[1, 1].imshow(np.abs(UMU), cmap="viridis")
    ax[1, 1].set_title("|U^T M U|")
    fig.colorbar(im, ax=ax[1, 1])
    fig.tight_layout()
    fig.savefig(config.diagnostics_viz, dpi=120)
    plt.close(fig)
