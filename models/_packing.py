"""Keep the user | item | brand tables in ONE contiguous HBM block.

The reference concatenates the three ``nn.Embedding`` weights on every forward (reference
``models/lightgcn.py:40``) and splits the result again (``:57-59``).  Here the three
``nn.Parameter``s are consecutive row-slices of a single [N, d] allocation, so the propagation
kernel reads layer 0 in place.  ``nn.Module.to()`` re-allocates every parameter separately
(reference ``main.py:467`` builds on CPU, then moves), so packing is (re)done lazily on the
first forward; the ``Parameter`` objects keep their identity, which is all ``optim.Adam``
(reference ``main.py:469``) and ``state_dict`` (``main.py:550,571``) rely on.
"""
import torch


def is_packed(params):
    p0 = params[0]
    end = p0.data_ptr()
    for p in params:
        if (not p.is_contiguous()) or p.data_ptr() != end or \
                p.untyped_storage().data_ptr() != p0.untyped_storage().data_ptr():
            return False
        end += p.numel() * p.element_size()
    return True


def pack_(params):
    """Re-home the parameters (in order) into one block; returns the [N,d] block view."""
    d = params[0].shape[1]
    n = sum(p.shape[0] for p in params)
    if not is_packed(params):
        block = torch.empty((n, d), dtype=params[0].dtype, device=params[0].device)
        r = 0
        with torch.no_grad():
            for p in params:
                block[r:r + p.shape[0]].copy_(p.data)
                p.data = block[r:r + p.shape[0]]
                r += p.shape[0]
    return params[0].data.as_strided((n, d), (d, 1))
