"""Integrated gradients over the node features and the adjacency values of a batch (kgcn/visualization.py:164-262).

The reference re-feeds the graph ``divide_number`` times with the perturbation targets scaled by ``(k + 1) /
divide_number`` (kgcn/feed.py:88-131: features and the first-channel adjacency VALUES are multiplied by the
scaling factor) and accumulates ``gradient * input / divide_number``.  Here each step is one forward + backward
through the C-ABI layers.  The gradient with respect to the adjacency values is the registered ``d values`` of
the batched SpMM ops (kgcn/bspmm_call.py:49-54 -> ``kgcn_bspmm_dvalues_f32``): pass the values as the tensors
inside the sparse triples handed to ``BatchedSpMM`` / ``BatchedConv`` / a plugin-mode ``GraphConv``.
"""
import torch

METHODS = ("ig", "grad_prod", "grad", "smooth_grad", "smooth_ig")


def integrated_gradients(score_fn, features, adj_values=None, divide_number=10, method="ig", noise_scale=0.1,
                         generator=None):
    """``score_fn(features, adj_values) -> scalar tensor`` (e.g. one class' prediction score summed over the batch).

    ``features``: the node-feature tensor; ``adj_values``: flat tensor of the (first-channel) adjacency values or
    ``None`` (the reference attributes first-channel values only, visualization.py:168).  Returns a dict with the
    attribution of every given input (same shape), plus ``sum_of_IG`` -- for ``method="ig"`` it approximates
    ``score(input) - score(0)`` (the reference's completeness check, visualization.py:254-257)."""
    if method not in METHODS:
        raise ValueError("method must be one of %s" % (METHODS,))
    inputs = {"features": features.detach()}
    if adj_values is not None:
        inputs["adjs"] = adj_values.detach()
    out = {k: torch.zeros_like(v) for k, v in inputs.items()}
    steps = 1 if method in ("grad_prod", "grad") else int(divide_number)
    for k in range(steps):
        scaling = (k + 1) / float(divide_number) if method in ("ig", "smooth_ig") else 1.0
        fed = {}
        for name, t in inputs.items():
            v = t * scaling
            if method in ("smooth_grad", "smooth_ig"):
                v = v + noise_scale * torch.randn(t.shape, device=t.device, dtype=t.dtype, generator=generator)
            fed[name] = v.requires_grad_(True)
        score = score_fn(fed["features"], fed.get("adjs"))
        grads = torch.autograd.grad(score, list(fed.values()), allow_unused=True)
        if "adjs" in fed and grads[list(fed).index("adjs")] is None:
            raise ValueError("score_fn did not use adj_values on the autograd tape: hand them to BatchedSpMM / BatchedConv or a "
                             "plugin-mode GraphConv (load_bspmm) as the `values` of the sparse triples")
        for (name, t), g in zip(inputs.items(), grads):
            if g is None:
                continue
            weight = t if method in ("ig", "grad_prod", "smooth_ig") else 1.0
            out[name] += g * weight / float(steps)
    out["sum_of_IG"] = float(sum(v.sum().item() for v in out.values()))
    return out
