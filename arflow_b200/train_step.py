"""One UFlow training step (model forward, UFlowLoss, backward, Adam) as a replayable unit.

Data-parallel modes (world_size > 1):
  * eager ("overlap"): a bucket of the flat gradient buffer is all-reduced on a side stream from a
    post-accumulate hook as soon as its last gradient is written, overlapping the rest of backward;
  * CUDA graph ("graph_split", the default with use_graph): forward+backward are one captured graph, the
    all-reduce of the whole flat buffer is ONE eager NCCL call between two graphs, Adam is the second graph.
    (Capturing the NCCL calls themselves inside the step graph hung on this pool's B200 boxes — torch 2.11 /
    NCCL 2.28.9, with and without the watchdog's async error handling — so the collective stays outside;
    22.9 MB over NVLink is ~0.1 ms against a ~14.5 ms step.)


Replaces, for the benchmark driver only, the per-step body of trainer/uflow_trainer.py:30-73 of the
reference: same maths (PWCFlow -> flows_fw/flows_bw -> cat -> UFlowLoss -> backward -> Adam with the
config's lr/betas/eps), restructured for the B200:
  * gradients live in ONE flat buffer (zeroed by one memset, all-reduced in a few large NCCL calls
    issued on a side stream as soon as a bucket's last gradient has been written, i.e. overlapped
    with the rest of backward),
  * no host synchronisation inside the step (the reference reads four `.item()`s and one level-
    dropout draw per level from the host), so the whole step is captured once in a CUDA graph and
    replayed: launch latency of the ~660 kernels of a step disappears.
"""
import torch
import torch.distributed as dist


class UFlowTrainStep:
    def __init__(self, model, loss_fn, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, use_graph=True, world_size=1,
                 n_buckets=3, global_census_norm=False, allreduce="auto"):
        if global_census_norm and world_size > 1:
            # batch-global census normaliser: one more collective inside the forward pass,
            # which rules out the captured fwd+bwd graph on this pool
            if use_graph:
                raise ValueError("global_census_norm needs use_graph=False (NCCL cannot be captured on this pool)")
            from . import uflow_utils
            uflow_utils.set_census_normaliser_group(dist.group.WORLD)
        self.model = model
        self.loss_fn = loss_fn
        self.world_size = world_size
        self.use_graph = use_graph
        self.allreduce_mode = None if world_size == 1 else "nccl"
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
            self.flat_grad = torch.zeros(total, device=dev, dtype=self.params[0].dtype)
            for p, (s0, e0) in zip(self.params, self._spans):
                # same strides as the parameter (dense, possibly channels-last): the fused Adam walks raw memory
                p.grad = self.flat_grad[s0:e0].as_strided(p.shape, p.stride())
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
            bounds = [int(round(total * k / n_buckets)) for k in range(n_buckets + 1)]
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
            self.reduced_log = []        # bucket ids in the order their all-reduce was issued (last step)
            self._hooks_on = not use_graph   # graph_split mode reduces the flat buffer in one call instead
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
            if not self._hooks_on:
                return
            if self._counts is None:     # discovery step
                self._seen.add(idx)
                return
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch_allreduce(b)
        return hook

    def _launch_allreduce(self, b):
        s, e = self._buckets[b]
        self.reduced_log.append(b)
        if self._on_cuda:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(self.flat_grad[s:e], op=dist.ReduceOp.AVG)
        else:                            # gloo (CPU tests): no AVG op
            dist.all_reduce(self.flat_grad[s:e])
            self.flat_grad[s:e] /= self.world_size

    def _finish_allreduce(self):
        if self._counts is None:
            # first step: reduce every bucket now and fix the per-bucket counts for the following steps
            for b in range(len(self._buckets)):
                self._launch_allreduce(b)
            self._counts = [sum(1 for i in self._seen if self._bucket_of[i] == b) for b in range(len(self._buckets))]
        if self._on_cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)

    # ---------------------------------------------------------------- the step
    def _fwd_bwd(self, img_pair):
        if self.world_size > 1:
            self.flat_grad.zero_()
            self.reduced_log = []
            if self._counts is not None:
                self._pending = list(self._counts)
        else:
            for p in self.params:
                p.grad = None
        res = self.model(img_pair, with_bk=True)
        flows = [torch.cat([a, b], 1) for a, b in zip(res['flows_fw'], res['flows_bw'])]
        loss, l_ph, l_sm, flow_mean, _ = self.loss_fn(flows, img_pair)
        loss.backward()
        return torch.stack([loss.detach(), l_ph.detach(), l_sm.detach(), flow_mean.detach()])

    def _step_impl(self, img_pair):
        out = self._fwd_bwd(img_pair)
        if self.world_size > 1:
            if self._hooks_on:
                self._finish_allreduce()
            else:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
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
        if self.world_size > 1:
            with torch.cuda.graph(self._graph):
                self._static_out = self._fwd_bwd(self._static_in)
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
        if self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
            self._graph_opt.replay()
        # a copy: the graph overwrites its static output on the next replay, a caller may keep the returned tensor
        return self._static_out.clone()
