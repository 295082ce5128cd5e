"""Drop-in for ``dddm/sampling.py``: Algorithm 2 with the fused bridge + update kernel (K3)."""
from __future__ import annotations

import weakref
from typing import Optional, Sequence, Tuple

import torch

from . import ops


@torch.no_grad()
def sample_dddm(
    model: torch.nn.Module,
    n_samples: int = 4096,
    steps: int = 20,
    eps_churn: float = 1.0,
    device: str = "cpu",
    data_shape: Sequence[int] | torch.Size | None = None,
    *,
    noise: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
    cuda_graph: bool = False,
    fused_noise: bool = True,
) -> torch.Tensor:
    """Algorithm 2 on the coarse grid t_0=0 < ... < t_N=1 — reference ``dddm/sampling.py:8-32``.

    Same signature, side effects (``model.to(device).eval()``) and RNG consumption order as the
    reference (x_T, then per step xi and z), so a seeded run draws the same noise as the reference
    on the same device.  ``noise=(x_T, xis, zs)`` (xis/zs indexed by the loop variable k) passes the
    noise in instead.  The bridge coefficients and the update run in ONE kernel per step; the
    times stay on the device (no per-step host synchronisation).  CUDA-only.

    ``fused_noise`` (default, only without ``noise``): the two Gaussian draws of a step (``sampling.py:27,30``) are
    generated INSIDE the update kernel — z in registers, the next step's xi written by the same launch — from the
    CUDA generator's Philox (seed, offset), bit-identical to the ``torch.randn_like`` calls they replace, and the
    generator is left at the same offset as after the reference's loop.

    ``cuda_graph=True`` (only without ``noise``) captures ONE step — the two noise draws, the backbone forward and
    the K3 update — as a CUDA graph and replays it ``steps`` times with (s, t) read from device memory: at 128
    samples per GPU (BASELINE config 5 on 8 GPUs) a step is ~100 small kernels and launch-bound otherwise.  The
    draw order is unchanged; the values come from the same Philox stream advanced through the graph.  The graph
    is cached per (model, batch shape, churn), so only the first call pays the capture.
    """
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("ddm_b200.sample_dddm runs on CUDA devices only (pass device='cuda'); "
                           "there is no CPU fallback")
    model = model.to(dev).eval()
    B = n_samples
    t_grid = torch.linspace(0.0, 1.0, steps + 1, device=dev)
    if data_shape is None:
        data_shape = (2,)
    if noise is None:
        x = torch.randn((B, *tuple(data_shape)), device=dev)
    else:
        x = noise[0].to(dev)
    if cuda_graph and noise is None and steps > 0:
        return _sample_graphed(model, x, t_grid, steps, float(eps_churn), fused_noise)
    if noise is None and fused_noise and steps > 0:
        return _sample_fused_noise(model, x, t_grid, steps, float(eps_churn))
    for k in reversed(range(steps)):
        s = t_grid[k:k + 1]
        t = t_grid[k + 1:k + 2]
        xi = torch.randn_like(x) if noise is None else noise[1][k].to(dev)
        xhat0 = model(x, t.repeat(B), xi)
        z = torch.randn_like(x) if noise is None else noise[2][k].to(dev)
        x = ops.bridge_step(x, xhat0.to(x.dtype), z, s, t, float(eps_churn))
    return x


def _generator(dev: torch.device) -> torch.Generator:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return torch.cuda.default_generators[idx]


def _sample_fused_noise(model, x: torch.Tensor, t_grid: torch.Tensor, steps: int, eps_churn: float) -> torch.Tensor:
    """The eager loop with both draws of a step inside K3.  Reference order per step: xi (offset o), denoiser,
    z (o + c), next xi (o + 2c), ...: the first xi comes from torch itself, every later one from the previous
    step's update launch at exactly the offset the reference would have drawn it."""
    dev, B = x.device, x.shape[0]
    gen = _generator(dev)
    seed = gen.initial_seed()
    c = ops.philox_increment(x.numel(), dev)
    x = x.contiguous().clone()
    xi = torch.randn_like(x)
    for k in reversed(range(steps)):
        s = t_grid[k:k + 1]
        t = t_grid[k + 1:k + 2]
        xhat0 = model(x, t.repeat(B), xi)
        off = gen.get_offset()  # where the reference would draw this step's z
        xi_next = torch.empty_like(x) if k > 0 else None
        ops.bridge_step_philox_(x, xhat0.to(x.dtype), xi_next, s, t, eps_churn, seed=seed, offset_z=off, offset_xi=off + c)
        gen.set_offset(off + (2 * c if k > 0 else c))
        xi = xi_next
    return x


_graph_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()  # model -> {(shape, churn, device): state}


def _sample_graphed(model, x: torch.Tensor, t_grid: torch.Tensor, steps: int, eps_churn: float,
                    fused_noise: bool = True) -> torch.Tensor:
    """Replays a cached one-step graph (captured on first use for this model / batch shape / churn: capture costs
    ~1 s, a replayed step ~1 ms).  The graph reads the model's parameters in place, so in-place weight updates are
    picked up; if parameters are re-allocated (``model.to(other_dtype)``) the key changes and a new graph is captured."""
    dev, B = x.device, x.shape[0]
    # the graph bakes in the addresses of the model's parameters and buffers: re-capture if any of them moved
    # (model.to(other dtype/device), load_state_dict(assign=True), ...)
    storage = hash(tuple(t.data_ptr() for t in list(model.parameters()) + list(model.buffers())))
    key = (tuple(x.shape), x.dtype, float(eps_churn), str(dev), storage, model.training, bool(fused_noise))
    per_model = _graph_cache.setdefault(model, {})
    st = per_model.get(key)
    if st is None:
        xs = x.clone()
        s_buf, t_buf = t_grid[:1].clone(), t_grid[:1].clone()
        xi_buf = torch.zeros_like(xs) if fused_noise else None
        ph_buf = torch.zeros(3, dtype=torch.int64, device=dev) if fused_noise else None

        def body():
            if fused_noise:  # xi_buf holds this step's xi; the update draws z and overwrites xi_buf with the next xi
                xhat0 = model(xs, t_buf.repeat(B), xi_buf)
                ops.bridge_step_philox_(xs, xhat0.to(xs.dtype), xi_buf, s_buf, t_buf, eps_churn, philox=ph_buf)
                return
            xi = torch.randn_like(xs)
            xhat0 = model(xs, t_buf.repeat(B), xi)
            z = torch.randn_like(xs)
            xs.copy_(ops.bridge_step(xs, xhat0.to(xs.dtype), z, s_buf, t_buf, eps_churn))

        rng = torch.cuda.get_rng_state(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()  # warm-up outside the capture (library plans, allocator); its draws are rolled back below
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        torch.cuda.set_rng_state(rng, dev)
        st = per_model[key] = (graph, xs, s_buf, t_buf, xi_buf, ph_buf)
    graph, xs, s_buf, t_buf, xi_buf, ph_buf = st
    xs.copy_(x)
    if fused_noise:
        # Philox positions of every step, computed on the host once: execution step j draws z at off0 + 2c*j and
        # the next xi at off0 + 2c*j + c (the reference's order: xi, z, xi, z, ...); the first xi is torch's own draw.
        gen = _generator(dev)
        seed = gen.initial_seed()
        seed = seed - (1 << 64) if seed >= (1 << 63) else seed  # two's complement into int64
        c = ops.philox_increment(x.numel(), dev)
        xi_buf.copy_(torch.randn_like(xs))
        off0 = gen.get_offset()
        table = torch.tensor([[seed, off0 + 2 * c * j, off0 + 2 * c * j + c] for j in range(steps)],
                             dtype=torch.int64).to(dev, non_blocking=True)
    for j, k in enumerate(reversed(range(steps))):
        s_buf.copy_(t_grid[k:k + 1])
        t_buf.copy_(t_grid[k + 1:k + 2])
        if fused_noise:
            ph_buf.copy_(table[j])
        graph.replay()
    if fused_noise:
        gen.set_offset(off0 + 2 * c * steps - c)  # the last launch's "next xi" is not part of the stream
    return xs.clone()


def clear_graph_cache() -> None:
    _graph_cache.clear()


def sample_dddm_chunked(model, n_samples: int, steps: int = 20, eps_churn: float = 1.0, device: str = "cuda",
                        data_shape=None, *, chunk: int = 4096, cuda_graph: bool = True, out_device=None) -> torch.Tensor:
    """Algorithm 2 for evaluation-sized sample counts (``eval_samples: 50000``, configs/cifar10_dit.yaml:32;
    ``train_cifar10_dit.py:325-335`` samples them in one call): ``n_samples`` drawn ``chunk`` at a time so that the
    backbone's activations stay bounded, every full chunk replaying ONE cached CUDA graph.  Chunks are concatenated on
    ``out_device`` (default: the sampling device).  Same per-chunk RNG order as :func:`sample_dddm`."""
    if n_samples < 0 or chunk < 1:
        raise ValueError("n_samples must be >= 0 and chunk >= 1")
    parts = []
    done = 0
    while done < n_samples:
        n = min(chunk, n_samples - done)
        x = sample_dddm(model, n, steps, eps_churn, device=device, data_shape=data_shape, cuda_graph=cuda_graph)
        parts.append(x if out_device is None else x.to(out_device))
        done += n
    if not parts:
        shape = (2,) if data_shape is None else tuple(data_shape)
        return torch.empty((0, *shape), device=out_device or device)
    return torch.cat(parts, dim=0)


def sample_dddm_sharded(model, n_samples: int, steps: int = 20, eps_churn: float = 1.0, data_shape=None, *,
                        group=None, gather: bool = True, seed: Optional[int] = None,
                        cuda_graph: bool = False) -> torch.Tensor:
    """Batch-sharded Algorithm 2 across the ranks of ``group`` (BASELINE config 5).

    Every rank samples ``n_samples / world`` images independently on its own GPU (no collective on
    the data path); one ``all_gather`` at the end assembles the full batch when ``gather`` is set.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if n_samples % world:
        raise ValueError(f"n_samples ({n_samples}) must be divisible by the number of ranks ({world})")
    dev = torch.device("cuda", torch.cuda.current_device())
    if seed is not None:
        torch.manual_seed(seed + rank)
    local = sample_dddm(model, n_samples // world, steps, eps_churn, device=str(dev), data_shape=data_shape,
                        cuda_graph=cuda_graph)
    if world == 1 or not gather:
        return local
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=0)
