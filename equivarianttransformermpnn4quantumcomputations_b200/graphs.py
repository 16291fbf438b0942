"""CUDA-graph replay of a training step.

A train step of this path is ~2 900 kernel launches of 2-500 us: enqueueing them from Python costs the host about as long
as the GPU needs to run them (bench.py reports both), so the step is captured once per problem signature
(atoms, edges, structures) and replayed.  Only the data-dependent head stays eager -- the neighbour list (its edge count
is read back to size the edge tensors) and the edge frames (`model.prepare`) -- and the optimizer update, which runs
eagerly on the gradients the replay leaves in place (its state and step count behave exactly as without graphs).
A batch with a new signature is captured again, up to `max_graphs` signatures (each capture keeps its activations in the
graphs' shared private memory pool: several GB for the OC20 step, ~100 GB for config 5 at 8 x 200 atoms); beyond that, or
with `max_graphs=0`, the step runs eagerly.

Real data has a new (atoms, edges) pair almost every step.  With `bucket=(n_mult, e_mult)` the prepared batch is padded
with one ghost structure to the next multiples (batching.pad_to_bucket: exact -- the ghost exchanges no messages with
real atoms and is masked out of the loss), so one capture serves every batch of its bucket; the loss must then weight
by `data["atom_mask"]` / `data["structure_mask"]` (batching.masked_mean)."""
import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, max_graphs=4, warmup=2, grad_sync=None, forward_loss=None, bucket=None):
        """model(data) -> outputs; loss_fn(outputs, data) -> scalar -- or `forward_loss(data)` -> scalar for steps that
        need more than that (MatPES: forces = -autograd.grad(E, pos, create_graph=True) inside the loss).
        `model.prepare(data)` must return the dict of data-dependent inputs the model then takes from `data` (OC20:
        edge_index / edge_distance / edge_distance_vec / edge_frames; MatPES v2: edge_index).  `grad_sync`: a callable
        (data parallel: parallel.GradientAllReducer.reduce, NCCL) that runs between the replay and the optimizer update,
        or an object with begin() / finish() (parallel.OverlappedGradientAllReducer) whose bucket all-reduces are issued
        from inside the backward pass and CAPTURED with it: the replayed graph then contains the exchange, overlapped
        with the rest of the backward pass."""
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.forward_loss = forward_loss if forward_loss is not None else (lambda d: loss_fn(model(d), d))
        self.overlapped = grad_sync if hasattr(grad_sync, "begin") else None
        self.grad_sync = None if self.overlapped is not None else grad_sync
        self.max_graphs, self.warmup = max_graphs, warmup
        self.bucket = bucket        # (atoms multiple, edges multiple) or None: exact signatures
        self.eager_steps = 0
        self.graphs = {}            # signature -> (graph, static inputs, static loss, gradient tensors, launches)
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        self.pool = None
        self.stream = None          # warm-up and every capture run on this one side stream (AccumulateGrad nodes bind to it)
        self.active = None          # signature whose gradient tensors are currently bound to the parameters
        self.replays = 0

    def _with_prepared(self, data):
        with torch.no_grad():
            prepared = self.model.prepare(data)
        out = dict(data)
        out.update(prepared)
        if self.bucket is not None:
            from . import batching
            out = batching.pad_to_bucket(out, *self.bucket)
        return out

    def _eager(self, full, sync=True):
        loss = self.forward_loss(full)
        self.optimizer.zero_grad(set_to_none=True)
        self._backward(loss, sync)
        return loss

    def _backward(self, loss, sync=True):
        if self.overlapped is not None and sync:
            self.overlapped.begin()
            loss.backward()
            self.overlapped.finish()
        else:
            loss.backward()

    def _capture(self, full, sig):
        static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in full.items()}
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        side = self.stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):               # warm-up off the capture: lazy tables, kernel attributes, allocator
            for _ in range(self.warmup):
                self._eager(static, sync=False)     # no exchange: ranks may capture at different steps
        torch.cuda.current_stream().wait_stream(side)
        self.optimizer.zero_grad(set_to_none=True)  # the captured backward then ASSIGNS the .grad tensors it allocates
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                    # the warm-up's activations sit in the caching allocator; the capture
                                                    # allocates its own copy in the graph pool -- do not hold both
        ops.reset_caches()
        from . import _lib
        n0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self.pool, stream=side):
            loss = self.forward_loss(static)
            self._backward(loss)
        ops.reset_caches()
        if self.pool is None:
            self.pool = graph.pool()
        self.graphs[sig] = (graph, static, loss, [p.grad for p in self.params], _lib.launch_count() - n0)
        self.active = sig

    def __call__(self, data):
        full = self._with_prepared(data)
        # host-side entries of the prepared inputs (e.g. the structure sizes the all-to-all attention loops over) are
        # baked into the capture: they are part of the signature
        sig = (int(full["pos"].shape[0]), int(full["edge_index"].shape[1]), len(full["natoms"])) + \
            tuple(v for k, v in sorted(full.items()) if isinstance(v, tuple))
        hit = self.graphs.get(sig)
        if hit is None:
            if len(self.graphs) >= self.max_graphs:
                self.eager_steps += 1
                loss = self._eager(full)
                self.active = None
                if self.grad_sync is not None:
                    self.grad_sync()
                self.optimizer.step()
                return loss
            self._capture(full, sig)
            hit = self.graphs[sig]
        graph, static, loss, grads, _ = hit
        for k, v in full.items():
            if torch.is_tensor(v):
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        self.replays += 1
        if self.active != sig:          # another signature (or an eager step) ran since: re-bind this graph's gradients
            for p, g in zip(self.params, grads):
                p.grad = g
            self.active = sig
        if self.grad_sync is not None:
            self.grad_sync()
        self.optimizer.step()
        return loss

    def release(self):
        """Drop every captured graph and the private memory pool they share (their activations stay allocated for as
        long as a graph exists); the gradients the last replay left in place go with them."""
        import gc
        self.optimizer.zero_grad(set_to_none=True)
        self.graphs.clear()
        self.pool = None
        self.active = None
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    def captured_launches(self, sig=None):
        """C-ABI kernel launches inside the captured step (the library's own count at capture time)."""
        hit = self.graphs.get(sig if sig is not None else self.active)
        return hit[4] if hit else 0
