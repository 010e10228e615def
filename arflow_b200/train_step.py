"""One UFlow training step (model forward, UFlowLoss, backward, Adam) as a replayable unit.

Data-parallel modes (world_size > 1), `allreduce=`:
  * "peer" (default on CUDA): gradients live in one flat buffer allocated by arflow_b200.comm.PeerAllReduce; a bucket
    is all-reduced by ONE kernel over NVLink peer memory (csrc/comm.cu) on a side stream as soon as its last gradient
    has been written, overlapping the rest of backward.  The kernel is an ordinary launch, so the whole step — forward,
    loss, backward, the bucket all-reduces, Adam — is captured in ONE CUDA graph; the batch-global census
    normaliser (a 16-byte all-reduce inside the forward pass) rides the same mechanism.
  * "nccl": the same buckets through torch.distributed.  Eager: overlapped on a side stream.  With use_graph:
    forward+backward are one captured graph, the all-reduce of the whole flat buffer is ONE eager NCCL call between
    two graphs, Adam is the second graph, no overlap (capturing the NCCL calls themselves inside the step graph hung
    on this pool's B200 boxes — torch 2.11 / NCCL 2.28.9, with and without the watchdog's async error handling).


Replaces, for the benchmark driver only, the per-step body of trainer/uflow_trainer.py:30-73 of the
reference: same maths (PWCFlow -> flows_fw/flows_bw -> cat -> UFlowLoss -> backward -> Adam with the
config's lr/betas/eps), restructured for the B200:
  * gradients live in ONE flat buffer (views with the parameters' own strides; autograd writes each gradient once, so
    there is no per-step memset or accumulate kernel), reduced bucket by bucket as described above,
  * no host synchronisation inside the step (the reference reads four `.item()`s and one level-dropout draw per level
    from the host), so the whole step is captured once in a CUDA graph and replayed: the launch latency of its ~340
    library kernels and the cuDNN / ATen launches between them disappears,
  * `__call__` returns a clone of the graph's static output tensor, so a caller may keep the returned losses.
"""
import torch
import torch.distributed as dist


class UFlowTrainStep:
    def __init__(self, model, loss_fn, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, use_graph=True, world_size=1,
                 n_buckets=4, global_census_norm=False, allreduce="auto", comm_ctas=32):
        dev0 = next(model.parameters()).device
        if allreduce == "auto":
            allreduce = "peer" if (world_size > 1 and dev0.type == "cuda") else "nccl"
        if allreduce == "fused":
            allreduce = "peer"
        if allreduce not in ("peer", "nccl"):
            raise ValueError("allreduce must be 'auto', 'peer' or 'nccl'")
        self._census_comm = None
        if global_census_norm and world_size > 1:
            from . import uflow_utils
            if allreduce == "peer":
                from .comm import PeerAllReduce
                self._census_comm = PeerAllReduce(4, dev0, ctas=1)
                uflow_utils.set_census_normaliser_group(self._census_comm)
            else:
                # batch-global census normaliser through NCCL: one more collective inside the forward pass, which
                # rules out the captured fwd+bwd graph on this pool
                if use_graph:
                    raise ValueError("global_census_norm with allreduce='nccl' needs use_graph=False "
                                     "(NCCL cannot be captured on this pool)")
                uflow_utils.set_census_normaliser_group(dist.group.WORLD)
        self.model = model
        self.loss_fn = loss_fn
        self.world_size = world_size
        self.use_graph = use_graph
        self.allreduce_mode = None if world_size == 1 else allreduce
        self._comm = None
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        if dev.type == "cuda" and getattr(model, "_nhwc", False):
            # channels-last conv stacks (fused_conv.py): keep the 4-D weights channels-last in place, so the per-step
            # `weight.contiguous(memory_format=channels_last)` of every layer is a no-op instead of a copy kernel
            # (state_dict, shapes and values are unchanged; gradients and Adam state follow the parameter's strides)
            for p in self.params:
                if p.dim() == 4:
                    p.data = p.data.contiguous(memory_format=torch.channels_last)
        # multi-GPU: flat gradient storage, parameter .grad tensors are views into it (16-byte aligned
        # spans so the accumulation kernels vectorise).  Single GPU: gradients are simply dropped to None
        # every step, autograd then hands its freshly written tensors over without an accumulate pass.
        self.flat_grad = None
        self._spans = []
        if world_size > 1:
            off = 0
            for p in self.params:
                n = p.numel()
                self._spans.append((off, off + n))
                off += (n + 3) // 4 * 4
            total = off
            if self.allreduce_mode == "peer":
                from .comm import PeerAllReduce
                self._comm = PeerAllReduce(total, dev, ctas=comm_ctas)
                self.flat_grad = self._comm.buffer[:total]
            else:
                self.flat_grad = torch.zeros(total, device=dev, dtype=self.params[0].dtype)
            # views with the parameter's own strides (dense, possibly channels-last): the fused Adam walks raw memory.
            # A step leaves .grad = None, autograd hands its freshly written gradient tensors over without an
            # accumulate pass, and a bucket's gradients are copied into their views by one multi-tensor copy when the
            # bucket is complete (no per-step memset of the flat buffer, no read-modify-write per parameter).
            self._views = [self.flat_grad[s0:e0].as_strided(p.shape, p.stride()) for p, (s0, e0) in zip(self.params, self._spans)]
        self.optimizer = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps,
                                          capturable=dev.type == "cuda", fused=dev.type == "cuda")
        # buckets over the flat buffer; a bucket is all-reduced (on a side stream, overlapping the rest of
        # backward) as soon as the last of its gradients has been written.  Parameters that never receive
        # a gradient (e.g. the unused context up-sampling layers of levels 0 and 1) are discovered on the
        # first step, which reduces everything after backward instead.
        self._buckets = []
        self._comm_stream = None
        self._on_cuda = dev.type == "cuda"
        if world_size > 1:
            if self._on_cuda:
                self._comm_stream = torch.cuda.Stream(device=dev)
            # buckets in the order backward completes them (the flat buffer starts with the decoders and ends with the
            # feature pyramid, whose gradients arrive last): the last bucket is small, because its all-reduce is the
            # only one nothing is left to overlap with
            fr = {1: [1.0], 2: [0.85, 1.0], 3: [0.5, 0.9, 1.0]}.get(n_buckets, None)
            if fr is None:
                fr = [0.9 * (k + 1) / (n_buckets - 1) for k in range(n_buckets - 1)] + [1.0]
            bounds = [0] + [int(round(total * f)) for f in fr]
            starts = [s0 for (s0, _) in self._spans]
            snapped = [0] + [min(starts, key=lambda s0: abs(s0 - b)) for b in bounds[1:-1]] + [total]
            snapped = sorted(set(snapped))
            self._buckets = [(snapped[i], snapped[i + 1]) for i in range(len(snapped) - 1)]
            self._bucket_of = []
            for (s, e) in self._spans:
                self._bucket_of.append(next(i for i, (bs, be) in enumerate(self._buckets) if bs <= s < be))
            self._counts = None          # per-bucket number of parameters that do get gradients
            self._seen = set()
            self._pending = [0] * len(self._buckets)
            self._arrived = [[] for _ in self._buckets]   # parameter indices whose gradient has been written, per bucket
            self.reduced_log = []        # bucket ids in the order their all-reduce was issued (last step)
            # graph_split (NCCL under use_graph) reduces the flat buffer in one call between two graphs instead
            self._hooks_overlap = not (use_graph and self.allreduce_mode == "nccl")
            for idx, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(self._make_hook(idx))
        self._graph_opt = None
        self._graph = None
        self.launches_per_step = None
        self._static_in = None
        self._static_out = None

    # ---------------------------------------------------------------- gradient all-reduce
    def _make_hook(self, idx):
        b = self._bucket_of[idx]

        def hook(_param):
            self._arrived[b].append(idx)
            if self._counts is None:     # discovery step
                self._seen.add(idx)
                return
            self._pending[b] -= 1
            if self._pending[b] == 0 and self._hooks_overlap:
                self._launch_allreduce(b)
        return hook

    def _gather_bucket(self, b):
        """Copy the bucket's freshly written gradients into their views of the flat buffer (one multi-tensor copy);
        views of parameters that received no gradient this step are zeroed."""
        got = self._arrived[b]
        if got:
            torch._foreach_copy_([self._views[i] for i in got], [self.params[i].grad for i in got])
        miss = [i for i in range(len(self.params)) if self._bucket_of[i] == b and i not in set(got)]
        if miss:
            torch._foreach_zero_([self._views[i] for i in miss])

    def _reduce_range(self, s, e):
        if self.allreduce_mode == "peer":
            # the range that ends the buffer is reduced after backward has finished: it may take the whole machine
            self._comm.all_reduce_(s, e, average=True, ctas=64 if e >= self.flat_grad.numel() else None)
        elif self._on_cuda:
            dist.all_reduce(self.flat_grad[s:e], op=dist.ReduceOp.AVG)
        else:                            # gloo (CPU tests): no AVG op
            dist.all_reduce(self.flat_grad[s:e])
            self.flat_grad[s:e] /= self.world_size

    def _launch_allreduce(self, b):
        s, e = self._buckets[b]
        self.reduced_log.append(b)
        if self._on_cuda:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                self._gather_bucket(b)
                self._reduce_range(s, e)
        else:
            self._gather_bucket(b)
            self._reduce_range(s, e)

    def _finish_allreduce(self):
        if self._counts is None:
            # first step: reduce every bucket now and fix the per-bucket counts for the following steps
            for b in range(len(self._buckets)):
                self._launch_allreduce(b)
            self._counts = [sum(1 for i in self._seen if self._bucket_of[i] == b) for b in range(len(self._buckets))]
        else:
            for b in range(len(self._buckets)):      # a bucket whose gradients did not all arrive (none expected)
                if b not in self.reduced_log:
                    self._launch_allreduce(b)
        if self._on_cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        for i, p in enumerate(self.params):          # the optimiser reads the reduced gradients from the flat buffer
            if p.grad is not None:
                p.grad = self._views[i]

    # ---------------------------------------------------------------- the step
    def _fwd_bwd(self, img_pair):
        for p in self.params:
            p.grad = None
        if self.world_size > 1:
            self.reduced_log = []
            self._arrived = [[] for _ in self._buckets]
            if self._counts is not None:
                self._pending = list(self._counts)
        res = self.model(img_pair, with_bk=True)
        flows = [torch.cat([a, b], 1) for a, b in zip(res['flows_fw'], res['flows_bw'])]
        loss, l_ph, l_sm, flow_mean, _ = self.loss_fn(flows, img_pair)
        loss.backward()
        return torch.stack([loss.detach(), l_ph.detach(), l_sm.detach(), flow_mean.detach()])

    def _gather_all(self):
        """graph_split: every gradient into the flat buffer, reduced by the caller in one NCCL call."""
        for b in range(len(self._buckets)):
            self._gather_bucket(b)

    def _assign_views(self):
        for i, p in enumerate(self.params):
            if p.grad is not None:
                p.grad = self._views[i]

    def _step_impl(self, img_pair):
        out = self._fwd_bwd(img_pair)
        if self.world_size > 1:
            if self._hooks_overlap:
                self._finish_allreduce()
            else:
                if self._counts is None:
                    self._counts = [sum(1 for i in self._seen if self._bucket_of[i] == b) for b in range(len(self._buckets))]
                self._gather_all()
                self._reduce_range(0, self.flat_grad.numel())
                self._assign_views()
        self.optimizer.step()
        return out

    def capture(self, example, warmup=3):
        """Warm up on a side stream (cuDNN autotune, Adam state, NCCL) and capture the step."""
        from . import _lib
        self._static_in = example.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_impl(self._static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        if self.world_size > 1 and not self._hooks_overlap:
            with torch.cuda.graph(self._graph):
                self._static_out = self._fwd_bwd(self._static_in)
                self._gather_all()
            self._assign_views()
            self._graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_opt):
                self.optimizer.step()
        else:
            with torch.cuda.graph(self._graph):
                self._static_out = self._step_impl(self._static_in)
        self.launches_per_step = _lib.launch_count() - n0   # arflow_b200 kernels inside one replay
        torch.cuda.synchronize()

    def __call__(self, img_pair):
        """img_pair: (B,6,H,W) device tensor.  Returns a 4-vector [loss, l_ph, l_sm, mean|flow|] on the device."""
        if not self.use_graph:
            return self._step_impl(img_pair)
        if self._graph is None:
            self.capture(img_pair)
        self._static_in.copy_(img_pair, non_blocking=True)
        self._graph.replay()
        if self._graph_opt is not None:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
            self._graph_opt.replay()
        # a copy: the graph overwrites its static output on the next replay, a caller may keep the returned tensor
        return self._static_out.clone()
