"""Shared gradient parity gate for the training-step tests (CPU estimate and GPU parity use the same rule)."""
import statistics

# bf16 GEMM operands / activations against the fp32 oracle: the CPU emulation of the kernels' rounding points
# (tests/test_train_oracle_cpu.py::test_host_orchestration_bf16_noise_estimate) measures <= 1.3e-2 per tensor.
GRAD_REL = 3e-2
# tensors whose true gradient is (numerically) zero - the key biases: softmax is invariant to a per-query shift - are compared
# against an absolute floor tied to the median gradient norm of the step instead (measured noise: 1e-3 of the median).
FLOOR_FRAC = 5e-3


def check_grads(grads, ref_grads, rel=GRAD_REL, floor_frac=FLOOR_FRAC):
    """grads / ref_grads: name -> tensor.  Returns the worst err / allowance ratio; asserts every tensor is within its allowance."""
    assert set(grads) == set(ref_grads), set(grads) ^ set(ref_grads)
    norms = {n: float(r.double().norm()) for n, r in ref_grads.items()}
    floor = floor_frac * statistics.median(norms.values())
    worst = 0.0
    for n, r in ref_grads.items():
        g = grads[n]
        assert g.shape == r.shape, (n, tuple(g.shape), tuple(r.shape))
        err = float((g.double().cpu() - r.double()).norm())
        allow = rel * norms[n] + floor
        assert err <= allow, f'{n}: |grad - ref| = {err:.3e} > {allow:.3e} (|ref| = {norms[n]:.3e})'
        worst = max(worst, err / allow)
    return worst
