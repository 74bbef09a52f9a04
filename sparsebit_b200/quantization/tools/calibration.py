"""Device-resident, streaming CalibrationRunner.

Same driver contract as sparsebit/quantization/tools/calibration.py:11-160
(``prepare_calibration()`` -> run the calibration batches through the model -> ``layerwise_calibration``)
and the same resulting qparams, but:

* the reference records the model inputs on the CPU, then replays the graph node by node, moving every
  activation of every batch host <-> device and caching it in full (calibration.py:38,137-160;
  observers/base.py:12-36).  Here the observers are fed *during* the calibration forwards by
  forward-pre-hooks: the streaming observers fold each batch into device-side running statistics
  (min/max keys, histograms, counts), so nothing is cached for MinMax / MovingAverage / ACIQ-gaus / KL
  and nothing ever leaves HBM;
* the layer-wise replay is kept for the two cases that need per-layer tensors -- AdaRound
  reconstruction and ``asym=True`` -- with the per-node storage held on the device and released as soon
  as a node's last user has run.

Under ``sparsebit_b200.distributed.enable(group)`` every rank streams its shard of the calibration set; the
runner then advances all quantizers' ``calc_qparams_steps`` in lockstep so that the statistics of the whole model
are merged with ONE packed MAX and ONE packed SUM all-reduce per round (``_finish_streaming``).
"""
import functools

import torch
import torch.fx as fx


def _is_quant_opr(module):
    return isinstance(module, torch.nn.Module) and hasattr(module, "input_quantizer") and hasattr(module, "weight_quantizer")


def _live(quantizer):
    return quantizer is not None and not quantizer.fake_fused


def trace_quant_model(model):
    """``torch.fx`` trace that keeps every quant operator a leaf (the role of the reference's QTracer,
    tools/graph_wrapper.py) -- needed only for the layer-wise replay."""

    class _Tracer(fx.Tracer):
        def is_leaf_module(self, m, qualname):
            return _is_quant_opr(m) or super().is_leaf_module(m, qualname)

    graph = _Tracer().trace(model)
    return fx.GraphModule(model, graph)


def _tensors_of(args):
    """Flattened tensor leaves of (nested) positional arguments, in order (calibration.py:24-31)."""
    out = []
    for a in args:
        if isinstance(a, torch.Tensor):
            out.append(a)
        elif isinstance(a, (list, tuple)):
            out.extend(_tensors_of(a))
    return out


class CalibrationRunner:
    def __init__(self, model, streaming=True, record_inputs=None):
        """``model``: an ``nn.Module`` whose quant operators expose ``input_quantizer`` /
        ``weight_quantizer`` / ``set_quant`` (a ``torch.fx.GraphModule`` for the layer-wise replay).
        ``streaming=False`` forces the replay for every configuration.  ``record_inputs`` (default:
        only when the model is a GraphModule) keeps references to the calibration inputs -- on the
        device, no copy -- so that ``layerwise_calibration`` may still choose the replay."""
        self.model = model
        self.streaming = streaming
        self.is_graph = isinstance(model, fx.GraphModule)
        self.record_inputs = self.is_graph if record_inputs is None else record_inputs
        assert streaming or self.is_graph, "the layer-wise replay needs a torch.fx.GraphModule (trace_quant_model)"
        self._handles = []
        self._inputs = {}     # placeholder name -> one entry per calibration batch
        self._order = []      # quant operators in first-execution order
        self._seen = set()

    # ------------------------------------------------------------------ pass 1: streaming hooks
    def prepare_calibration(self):
        self.model.eval()  # the reference's replay runs every operator in eval mode (calibration.py:141-142)
        seen = set()
        for module in self.model.modules():
            if _is_quant_opr(module) and id(module) not in seen:
                seen.add(id(module))
                self._handles.append(module.register_forward_pre_hook(self._feed_observer))
        if self.record_inputs:
            self._hook_graph_inputs()
        self.builder = self  # the reference asserts on this attribute (calibration.py:72)

    def _hook_graph_inputs(self):
        """Keep references to the calibration inputs for a possible replay.  Like the reference
        (calibration.py:16-62) each graph input is captured at the first operator that consumes it -- its
        QuantModel calls ``self.model.forward(...)`` directly (quant_model.py:206-207), which bypasses hooks
        on the graph module itself; inputs consumed first by a function / method node fall back to a hook on
        the graph module."""
        placeholders = [n for n in self.model.graph.nodes if n.op == "placeholder"]
        self._inputs = {p.name: [] for p in placeholders}
        pending = set(placeholders)
        for node in self.model.graph.nodes:
            if node.op != "call_module":
                continue
            wanted = [(pos, a.name) for pos, a in enumerate(node.args) if isinstance(a, fx.Node) and a in pending]
            if wanted:
                module = self.model.get_submodule(node.target)
                self._handles.append(module.register_forward_pre_hook(functools.partial(self._record_at_consumer, wanted=wanted)))
                pending -= {a for a in node.args if isinstance(a, fx.Node)}
        if pending:
            order = [p.name for p in placeholders]
            names = {p.name for p in pending}
            self._handles.append(self.model.register_forward_pre_hook(
                lambda module, args: [self._inputs[n].append(_detached(a)) for n, a in zip(order, args) if n in names] and None))

    def _record_at_consumer(self, module, args, wanted):
        for pos, name in wanted:
            self._inputs[name].append(_detached(args[pos]))

    def _feed_observer(self, module, args):
        if id(module) not in self._seen:
            self._seen.add(id(module))
            self._order.append(module)
        if self.streaming and _live(module.input_quantizer):
            q = module.input_quantizer
            # a retained batch may alias the live activation unless this operator overwrites its input
            kw = {"alias_ok": not getattr(module, "inplace", False)} if hasattr(q, "calc_qparams_steps") else {}
            for x in _tensors_of(args):
                q.update_observer(x.detach(), **kw)

    def _remove_hooks(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    # ------------------------------------------------------------------ pass 2: qparams
    def layerwise_calibration(self, device=None, asym=False, w_quant=False, a_quant=False):
        """calibration.py:64-98.  ``device`` is accepted for signature parity; tensors stay where the
        calibration forwards produced them."""
        assert hasattr(self, "builder"), "run self.prepare_calibration first!"
        self._remove_hooks()
        wants_replay = asym or any(
            m.weight_quantizer is not None and m.weight_quantizer.TYPE.lower() == "adaround" for m in self._quant_oprs())
        if self.streaming and not wants_replay:
            self._finish_streaming()
        else:
            assert self.is_graph and self._inputs and all(self._inputs.values()), \
                "AdaRound / asym calibration replays the graph: pass a GraphModule and run the calibration batches"
            if self.streaming:  # statistics gathered in pass 1 are recomputed per layer
                for m in self._quant_oprs():
                    if _live(m.input_quantizer):
                        m.input_quantizer.observer.data_cache.reset()
            self._replay(asym, w_quant, a_quant)
        self._inputs = {}

    def _quant_oprs(self):
        if self._order:
            return list(self._order)
        return [m for m in self.model.modules() if _is_quant_opr(m)]

    def _finish_streaming(self):
        """qparams of EVERY quantizer of the model in one lockstep sweep: with sharded calibration
        (sparsebit_b200.distributed.enable) the statistics of all observers cross the ranks in ONE packed MAX and
        ONE packed SUM all-reduce per round (2 collectives for MinMax / MSE / KL / ACIQ models, 4 with Percentile)
        instead of one to five per quantizer."""
        from ... import distributed as sbdist

        feature, todo = [], []
        for module in self._quant_oprs():
            if _live(module.input_quantizer):
                feature.append(module.input_quantizer)
                todo.append(module.input_quantizer)
            wq = module.weight_quantizer
            if wq is not None:
                wq.update_observer(module.weight, **({"alias_ok": True} if hasattr(wq, "calc_qparams_steps") else {}))
                todo.append(wq)
        sbdist.drive_all([_qparams_steps(q) for q in todo])
        for q in feature:
            q.observer.data_cache.reset()

    @staticmethod
    def _calibrate_weight(module):
        wq = module.weight_quantizer
        if wq is None:
            return
        wq.update_observer(module.weight, **({"alias_ok": True} if hasattr(wq, "calc_qparams_steps") else {}))
        wq.calc_qparams()

    # ------------------------------------------------------------------ layer-wise replay (device resident)
    def _replay(self, asym, w_quant, a_quant):
        from ..quantizers.adaround import reconstruct_qlayer

        graph = self.model.graph
        n_batches = min(len(v) for v in self._inputs.values())
        float_env, quant_env = {}, {}
        pending = {n: len(n.users) for n in graph.nodes}

        def release(node):
            for inp in node.all_input_nodes:
                pending[inp] -= 1
                if pending[inp] == 0:
                    float_env.pop(inp, None)
                    quant_env.pop(inp, None)

        for node in graph.nodes:
            if node.op == "placeholder":
                float_env[node] = self._inputs[node.name][:n_batches]
                if asym:
                    quant_env[node] = float_env[node]
                continue
            if node.op == "output":
                continue
            module = self.model.get_submodule(node.target) if node.op == "call_module" else None
            quant_opr = module if _is_quant_opr(module) else None
            # input quantizer: always on the float activations (calibration.py:100-115)
            if quant_opr is not None and _live(quant_opr.input_quantizer):
                for inp in node.all_input_nodes:
                    for x in float_env[inp]:
                        if isinstance(x, torch.Tensor):
                            # stored activations are never written to (in-place operators get private copies)
                            q = quant_opr.input_quantizer
                            q.update_observer(x, **({"alias_ok": True} if hasattr(q, "calc_qparams_steps") else {}))
                quant_opr.input_quantizer.calc_qparams()
                quant_opr.input_quantizer.observer.data_cache.reset()
            float_env[node] = self._run_node(node, module, float_env, n_batches)
            if quant_opr is not None:
                quant_opr.set_quant(w_quant=False, a_quant=False)
            if quant_opr is not None and quant_opr.weight_quantizer is not None:
                self._calibrate_weight(quant_opr)
                if quant_opr.weight_quantizer.TYPE.lower() == "adaround":
                    assert len(node.all_input_nodes) == 1, "AdaRound not supports the oprs which has more than one inputs"
                    source = quant_env if asym else float_env
                    reconstruct_qlayer(quant_opr, torch.cat(source[node.all_input_nodes[0]], dim=0),
                                       torch.cat(float_env[node], dim=0), a_quant=a_quant,
                                       **getattr(self, "adaround_kwargs", {}))
            if asym:
                if quant_opr is not None:
                    quant_opr.set_quant(w_quant, a_quant)
                quant_env[node] = self._run_node(node, module, quant_env, n_batches)
                if quant_opr is not None:
                    quant_opr.set_quant(w_quant=False, a_quant=False)
            release(node)

    def _run_node(self, node, module, env, n_batches):
        if node.op == "call_module":
            module.eval()
        outs = []
        with torch.no_grad():
            for b in range(n_batches):
                if node.op == "get_attr":
                    outs.append(_fetch_attr(self.model, node.target))
                    continue
                args = fx.node.map_arg(node.args, lambda n: env[n][b])
                kwargs = fx.node.map_arg(node.kwargs, lambda n: env[n][b])
                if getattr(module, "inplace", False) or kwargs.get("inplace", False) or (
                        node.op == "call_method" and node.target.endswith("_")):
                    # the stored activations are shared with the node's siblings and with the other pass
                    # (float / quantised): an in-place operator must work on a private copy
                    args = tuple(a.clone() if isinstance(a, torch.Tensor) else a for a in args)
                if node.op == "call_module":
                    outs.append(module(*args, **kwargs))
                elif node.op == "call_function":
                    outs.append(node.target(*args, **kwargs))
                elif node.op == "call_method":
                    outs.append(getattr(args[0], node.target)(*args[1:], **kwargs))
                else:
                    raise NotImplementedError(node.op)
        return outs


def _qparams_steps(q):
    """``calc_qparams`` of any quantizer as a step generator.  Native quantizers provide it.  A quantizer of the
    unmodified reference (``install()`` mode) that uses the stock ``Quantizer.calc_qparams``
    (quantizers/base.py:33-39) on top of a native observer is restated here so that its observer's statistics
    still join the packed collectives; anything else runs as one opaque step."""
    if hasattr(q, "calc_qparams_steps"):
        return q.calc_qparams_steps()
    stock = False
    for klass in type(q).__mro__:
        if "calc_qparams" in vars(klass):
            stock = klass.__module__ == "sparsebit.quantization.quantizers.base"
            break

    def steps():
        if stock and hasattr(q.observer, "calc_qparams_steps"):
            if q.fake_fused:
                return q.scale, q.zero_point
            scale, zero_point = yield from q.observer.calc_qparams_steps()
            q.scale = q._broadcast_qparams(scale)
            q.zero_point = q._broadcast_qparams(zero_point)
            return q.scale, q.zero_point
        return q.calc_qparams()
        yield  # pragma: no cover  (makes this a generator)

    return steps()


def _detached(x):
    return x.detach() if isinstance(x, torch.Tensor) else x


def _fetch_attr(root, target):
    obj = root
    for part in target.split("."):
        obj = getattr(obj, part)
    return obj
