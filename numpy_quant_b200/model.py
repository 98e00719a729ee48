"""Graph import + float / quantized executors with the public API of the reference's
`numpy_quant/model.py` (`Model.from_onnx`, `model(inputs, profile)`,
`model.quantize(calibration_inputs, bit_width)`, `QModel.__call__`), re-designed so that
every value stays resident in HBM between nodes and each node is one (or zero) kernel
launch on the current CUDA stream.

Differences to the reference that a caller can observe are limited to keyword-only
extras: `retain=False` (free intermediates after their last consumer and fuse multi-node
chains such as GELU) and `group=` / torch.distributed for sharded calibration.
"""
from __future__ import annotations

from collections import OrderedDict
from time import time
from typing import Any, List, Optional, Union

import numpy as np
import torch

from . import _lib, kernels as K
from . import onnx_lite
from .numpy_quantization import quant_parameters
from .tensor import (FTensor, ITensor, QTensor, Tensor, concat, fconv2d, qconv2d, qtensor_from_operand, quantize_tensor_nhwc,
                     quantize_tensor_patches,
                     quantize_tensor, where, _to_device)


class Constant:
    def __init__(self, name: str, outputs: List["Node"], data: Tensor = None):
        self.name = name
        self.outputs = outputs
        self.data = data

    def __repr__(self):
        return f"Constant({self.name})"


class Variable:
    def __init__(self, name: str, inputs: List["Node"], outputs: List["Node"], data: Tensor = None):
        self.name = name
        self.inputs = inputs
        self.outputs = outputs
        self.data = data

    def __repr__(self):
        return f"Variable({self.name})"


Value = Union[Constant, Variable]


class Node:
    def __init__(self, name: str, op: str, attrs: dict, inputs: List[Value], outputs: List[Value]):
        self.name = name
        self.op = op
        self.attrs = attrs
        self.inputs = inputs
        self.outputs = outputs

    def __repr__(self):
        return f"Node({self.name})"


def _onnx_api(onnx_model):
    """Helpers matching the proto flavour: the real `onnx` package if the model came from it,
    else the dependency-free reader."""
    if isinstance(onnx_model, onnx_lite.ModelProto):
        return onnx_lite.to_array, onnx_lite.get_attribute_value, onnx_lite.TensorProto
    import onnx
    import onnx.numpy_helper
    return onnx.numpy_helper.to_array, onnx.helper.get_attribute_value, onnx.TensorProto


# --------------------------------------------------------------------------------------
# operator table (reference model.py:65-213)
# --------------------------------------------------------------------------------------
def _op_constant(inputs, attrs):
    value = attrs["value"]
    if value.dtype == np.float32:
        return [FTensor(value)]
    if value.dtype == np.int64:
        return [ITensor(value)]
    cls = value.dtype.__class__
    raise ValueError(f"Constant value type {cls.__module__}.{cls.__qualname__} not supported.")


def _op_constant_of_shape(inputs, attrs):
    value = attrs["value"]
    y = np.full(tuple(inputs[0].data), fill_value=value, dtype=value.dtype)
    return _op_constant([], {"value": y})


def _op_conv(inputs, attrs):
    x, w, b = inputs
    if isinstance(x, QTensor) and isinstance(w, QTensor):
        return [qconv2d(x, w, b, tuple(attrs["pads"]), tuple(attrs["strides"]))]
    return [fconv2d(x, w, b, tuple(attrs["pads"]), tuple(attrs["strides"]))]


def _op_gemm(inputs, attrs):
    x, w, b = inputs
    if attrs.get("transA"):
        x = x.T
    if attrs.get("transB"):
        w = w.T
    return [x.matmul(w) + b]


def _op_layernorm(inputs, attrs):
    x, scale, bias = inputs
    nd = len(x.shape.data)
    if attrs["axis"] % nd == nd - 1:
        return [x.layernorm(scale, bias, attrs["epsilon"])]
    mean = x.mean(axis=attrs["axis"], keepdims=True)          # generic recipe, model.py:141-150
    d = x + (-mean)
    var = (d * d).mean(axis=attrs["axis"], keepdims=True)
    normalized = d * (var + attrs["epsilon"]).sqrt().inv()
    return [normalized * scale + bias]


def _op_slice(inputs, attrs):
    x = inputs[0]
    starts, ends, axes = inputs[1].data, inputs[2].data, inputs[3].data
    slices = [slice(None, None, None)] * x.shape.size
    for s, e, a in zip(starts, ends, axes):
        slices[a] = slice(int(s), int(e))
    return [x.__getitem__(tuple(slices))]


_OPS = {
    "Add": lambda i, a: [i[0] + i[1]],
    "Concat": lambda i, a: [concat(list(i), axis=a["axis"])],
    "Constant": _op_constant,
    "ConstantOfShape": _op_constant_of_shape,
    "Conv": _op_conv,
    "Div": lambda i, a: [i[0].div(i[1])],
    "Equal": lambda i, a: [i[0] == i[1]],
    "Erf": lambda i, a: [i[0].erf()],
    "Expand": lambda i, a: [i[0].expand(i[1])],
    "Gather": lambda i, a: [i[0].take(i[1], axis=a["axis"])],
    "Gemm": _op_gemm,
    "Identity": lambda i, a: [i[0].copy()],
    "LayerNormalization": _op_layernorm,
    "MatMul": lambda i, a: [i[0].matmul(i[1])],
    "Mul": lambda i, a: [i[0] * i[1]],
    "ReduceMean": lambda i, a: [i[0].mean(a["axis"] if "axis" in a else a["axes"][0], keepdims=bool(a.get("keepdims", 1)))],
    "Relu": lambda i, a: [i[0].relu()],
    "Reshape": lambda i, a: [i[0].reshape(i[1])],
    "Sigmoid": lambda i, a: [i[0].sigmoid()],
    "Shape": lambda i, a: [i[0].shape if not isinstance(i[0], QTensor) else ITensor(np.array(i[0].shape, np.int64))],
    "Slice": _op_slice,
    "Softmax": lambda i, a: [i[0].softmax(axis=a["axis"])],
    "Tanh": lambda i, a: [i[0].tanh()],
    "Transpose": lambda i, a: [i[0].transpose(a["perm"])],
    "Unsqueeze": lambda i, a: [i[0].expand_dims(axis=i[1])],
    "Where": lambda i, a: [where(i[0], i[1], i[2])],
}


def onnx_operator_implementation(op: str, inputs: list, attrs: dict) -> list:
    try:
        fn = _OPS[op]
    except KeyError:
        raise ValueError(f"ONNX operand {op} not supported.") from None
    return fn(inputs, attrs)


# --------------------------------------------------------------------------------------
# multi-node fusion patterns (used by retain=False)
# --------------------------------------------------------------------------------------
def _scalar_const(value: Value, producers: dict) -> Optional[float]:
    node = producers.get(value.name)
    if node is None or node.op != "Constant":
        return None
    v = node.attrs["value"]
    if v.dtype != np.float32 or v.size != 1:
        return None
    return float(v.reshape(-1)[0])


def find_gelu_chains(nodes: List[Node]) -> dict:
    """Div(x,c1) -> Erf -> Add(.,c2) -> Mul(x,.) -> Mul(.,c3) with single-consumer intermediates.
    Returns {name of the Div node: (x value, c1, c2, c3, [names of the 3 middle nodes], name of
    the last Mul node)}: the fused kernel runs at the Div node (x is alive there) and its result
    is handed to the output of the last node."""
    producers = {o.name: n for n in nodes for o in n.outputs}
    found = {}
    for n in nodes:
        if n.op != "Div" or len(n.outputs[0].outputs) != 1:
            continue
        c1 = _scalar_const(n.inputs[1], producers)
        x = n.inputs[0]
        erf = n.outputs[0].outputs[0]
        if c1 is None or erf.op != "Erf" or len(erf.outputs[0].outputs) != 1:
            continue
        add = erf.outputs[0].outputs[0]
        if add.op != "Add" or len(add.outputs[0].outputs) != 1:
            continue
        oth = [v for v in add.inputs if v is not erf.outputs[0]]
        c2 = _scalar_const(oth[0], producers) if len(oth) == 1 else None
        mul = add.outputs[0].outputs[0]
        if c2 is None or mul.op != "Mul" or len(mul.outputs[0].outputs) != 1:
            continue
        if not any(v is x for v in mul.inputs):
            continue
        mul2 = mul.outputs[0].outputs[0]
        if mul2.op != "Mul":
            continue
        oth = [v for v in mul2.inputs if v is not mul.outputs[0]]
        c3 = _scalar_const(oth[0], producers) if len(oth) == 1 else None
        if c3 is None:
            continue
        found[n.name] = (x, c1, c2, c3, [erf.name, add.name, mul.name], mul2.name)
    return found


class Model:
    def __init__(self, nodes: list, values: list, inputs: List[Variable], outputs: List[Variable]):
        self.nodes = nodes
        self.values = values
        self.inputs = inputs
        self.outputs = outputs

    def __repr__(self):
        return f"Model(nodes={self.nodes}, values={self.values}, inputs={self.inputs}, outputs={self.values})"

    def __str__(self):
        res = "Model(\n"
        for k, v in self.__dict__.items():
            if not isinstance(v, list):
                continue
            res += f"  {k}=[\n"
            for e in v:
                res += f"    {e}\n"
            res += "  ],\n"
        res += ")\n"
        return res

    def __del__(self):
        # break node <-> value reference cycles so device buffers are released by refcount
        # (reference model.py:236-247; test_delete.py relies on it)
        for node in getattr(self, "nodes", []):
            node.inputs = []
            node.outputs = []
        for value in getattr(self, "values", []):
            if isinstance(value, Variable):
                value.inputs = []
            value.outputs = []

    def release(self) -> None:
        """Drop every intermediate held on `Variable.data` (frees HBM; constants stay)."""
        for value in self.values:
            if isinstance(value, Variable) and not (value.inputs and value.inputs[0].op == "Constant"):
                value.data = None           # outputs of Constant nodes are immutable uploads: kept

    @classmethod
    def from_onnx(cls, onnx_model):
        """Build the graph from an `onnx.ModelProto` or an `onnx_lite.ModelProto`
        (anything exposing .graph.{initializer,input,node,output}); reference model.py:249-292.
        Initializers are uploaded to HBM once, here."""
        to_array, get_attr, tensor_type = _onnx_api(onnx_model)
        graph = onnx_model.graph
        value_dict: dict = {}
        for t in graph.initializer:
            arr = np.array(to_array(t))
            value_dict[t.name] = Constant(t.name, outputs=[], data=FTensor(arr))
        inputs: List[Value] = []
        for t in graph.input:
            if t.name in value_dict and isinstance(value_dict[t.name], Constant):
                continue                                   # initializer also listed as input (old exporters)
            value_dict[t.name] = Variable(t.name, inputs=[], outputs=[])
            inputs.append(value_dict[t.name])
        nodes: "OrderedDict[str, Node]" = OrderedDict()
        for n in graph.node:
            attrs = {}
            for a in n.attribute:
                v = get_attr(a)
                attrs[a.name] = to_array(v) if isinstance(v, tensor_type) else v
            node = Node(n.name, n.op_type, attrs, [], [])
            for name in n.input:
                if name not in value_dict:
                    value_dict[name] = Variable(name, inputs=[], outputs=[])
                value_dict[name].outputs.append(node)
            for name in n.output:
                if name not in value_dict:
                    value_dict[name] = Variable(name, inputs=[node], outputs=[])
                else:
                    value_dict[name].inputs.append(node)
            node.inputs = [value_dict[name] for name in n.input]
            node.outputs = [value_dict[name] for name in n.output]
            nodes[n.name or f"node_{len(nodes)}"] = node
        outputs = [value_dict[t.name] for t in graph.output]
        return cls(list(nodes.values()), list(value_dict.values()), inputs, outputs)

    # ---------------------------------------------------------------- float executor
    def __call__(self, inputs: List[np.ndarray], profile=False):
        """Float32 interpreter (reference model.py:294-326); values stay in HBM."""
        time_per_op_types = {op: 0.0 for op in {n.op for n in self.nodes}}
        for array, variable in zip(inputs, self.inputs):
            if isinstance(array, torch.Tensor):
                variable.data = FTensor(array)
            elif array.dtype == np.float32:
                variable.data = FTensor(array)             # upload == the reference's defensive copy
            elif array.dtype == np.int64:
                variable.data = ITensor(array.copy())
            else:
                raise ValueError(f"Array dtype {array.dtype} not supported")
        for node in self.nodes:
            node_inputs = [i.data for i in node.inputs]
            if profile:
                torch.cuda.synchronize()
                stime = time()
            outputs = onnx_operator_implementation(node.op, node_inputs, node.attrs)
            if profile:
                torch.cuda.synchronize()
                time_per_op_types[node.op] += time() - stime
            for o, tensor in zip(node.outputs, outputs):
                o.data = tensor
        output_tensors = [out_var.data.data for out_var in self.outputs]
        if profile:
            return output_tensors, time_per_op_types
        return output_tensors

    # ---------------------------------------------------------------- calibration
    def _value_min_max(self, group=None):
        """Global min / max of every value after a float pass (reference model.py:332-336): one
        reduction kernel per float value into a single stats buffer, one D2H copy; when the
        calibration batch is sharded over ranks the buffer is all-reduced (max of [max, -min]),
        which is exact, so every rank derives bit-identical quantization parameters."""
        fvals = [v for v in self.values if isinstance(v.data, FTensor)]
        vmin, vmax = {}, {}
        if fvals:
            mm = K.minmax_slots(len(fvals), fvals[0].data.device_tensor.device)
            for slot, v in enumerate(fvals):
                K.minmax_into(v.data.device_tensor, mm, slot)
            from .distributed import allreduce_minmax
            stats = allreduce_minmax(mm, group).cpu().numpy()
            for slot, v in enumerate(fvals):
                vmin[v.name], vmax[v.name] = np.float32(stats[slot, 0]), np.float32(stats[slot, 1])
        for v in self.values:
            if v.name in vmin:
                continue
            data = v.data.data                              # ITensor (host int64) statistics
            flat = data.reshape((data.shape[0], -1) if data.shape else (-1,))
            vmin[v.name], vmax[v.name] = np.mean(flat.min()), np.mean(flat.max())
        return vmin, vmax

    def _calibrate_streaming(self, inputs: list, group=None):
        """The calibration pass of `quantize(keep_values=False)`: the float forward of `__call__` with the min / max
        of every float value reduced as soon as it is produced and the activation freed after its last consumer, so
        the pass holds the live set instead of all 700+ intermediate tensors (at ViT-B batch 512 that is a few GB
        instead of more than the 180 GB of HBM).  Same statistics, same all-reduce as `_value_min_max`."""
        for array, variable in zip(inputs, self.inputs):
            if isinstance(array, torch.Tensor) or array.dtype == np.float32:
                variable.data = FTensor(array)
            elif array.dtype == np.int64:
                variable.data = ITensor(array.copy())
            else:
                raise ValueError(f"Array dtype {array.dtype} not supported")
        slot = {v.name: i for i, v in enumerate(self.values)}
        dev = next(v.data.device_tensor.device for v in self.values if isinstance(v.data, FTensor))
        mm = K.minmax_slots(len(self.values), dev)
        is_float = set()

        def record(v):
            if isinstance(v.data, FTensor) and v.name not in is_float:
                K.minmax_into(v.data.device_tensor, mm, slot[v.name])
                is_float.add(v.name)

        for v in self.values:
            if v.data is not None:
                record(v)                                   # constants and inputs
        keep = {id(v) for v in self.outputs} | {id(v) for v in self.inputs}
        remaining = {id(v): len({id(n) for n in v.outputs}) for v in self.values}
        for node in self.nodes:
            outputs = onnx_operator_implementation(node.op, [i.data for i in node.inputs], node.attrs)
            for o, tensor in zip(node.outputs, outputs):
                o.data = tensor
                record(o)
            for i in {id(v): v for v in node.inputs}.values():
                remaining[id(i)] -= 1
                if remaining[id(i)] == 0 and isinstance(i, Variable) and id(i) not in keep and isinstance(i.data, FTensor):
                    i.data = None                           # statistics taken, no consumer left
        from .distributed import allreduce_minmax
        stats = allreduce_minmax(mm, group).cpu().numpy()
        vmin, vmax = {}, {}
        for v in self.values:
            if v.name in is_float:
                vmin[v.name], vmax[v.name] = np.float32(stats[slot[v.name], 0]), np.float32(stats[slot[v.name], 1])
            else:
                data = v.data.data                          # ITensor (host int64) statistics
                flat = data.reshape((data.shape[0], -1) if data.shape else (-1,))
                vmin[v.name], vmax[v.name] = np.mean(flat.min()), np.mean(flat.max())
        return vmin, vmax

    def quantize(self, calibration_inputs: list, bit_width=8, *, group=None, keep_values: bool = True):
        """Calibrate on `calibration_inputs` and return the quantized model
        (reference model.py:328-442): per-tensor affine, min/max calibrated; constants
        symmetric `bit_width` bits, activations asymmetric, Gemm / Add biases 4*bit_width bits.
        keep_values=False frees every float activation of the calibration pass after its last use (`Value.data` of
        intermediates is then not readable afterwards, unlike the reference) -- same parameters, a fraction of the HBM."""
        if not 2 <= int(bit_width) <= 8:
            raise ValueError("bit_width must be in 2..8 on the B200 path (int8 tensor-core operands)")
        node_dict = {node.name: node for node in self.nodes}
        value_dict = {value.name: value for value in self.values}
        if keep_values:
            self(calibration_inputs)
            value_min_dict, value_max_dict = self._value_min_max(group)
        else:
            value_min_dict, value_max_dict = self._calibrate_streaming(calibration_inputs, group)

        def get_quantization_params(value: Value, asymmetric: bool):
            scale, zero_point = quant_parameters(value_min_dict[value.name], value_max_dict[value.name],
                                                 bit_width=bit_width, asymmetric=asymmetric)
            return QuantizationParams(scale, zero_point)

        qnodes_dict: "OrderedDict[str, Node]" = OrderedDict()
        qvalues_dict: dict = {}
        qparams_per_value: dict = {}
        for value in self.inputs:
            qvalues_dict[value.name] = value                # shared with the float model (model.py:350)
            qparams_per_value[value.name] = get_quantization_params(value, isinstance(value, Variable))
        for value in self.values:
            if isinstance(value, Constant):
                qp = get_quantization_params(value, False)
                qvalues_dict[value.name] = Constant(value.name, [],
                                                    quantize_tensor(value.data, bit_width, qp.scale, qp.zero_point))
                qparams_per_value[value.name] = qp
        for node in self.nodes:
            out_val = node.outputs[0]
            if node.op == "Gemm":
                for input_value in node.inputs[:2]:
                    if isinstance(input_value, Variable):
                        qvalues_dict[input_value.name] = Variable(input_value.name, [], [], None)
                        qparams_per_value[input_value.name] = get_quantization_params(input_value, True)
                bias = node.inputs[2]
                bias_scale = qparams_per_value[node.inputs[0].name].scale * qparams_per_value[node.inputs[1].name].scale
                qparams_per_value[bias.name] = QuantizationParams(bias_scale, None)
                qvalues_dict[bias.name] = Constant(bias.name, [], quantize_tensor(bias.data, 4 * bit_width, bias_scale, None))
                qparams_per_value[out_val.name] = get_quantization_params(out_val, True)
            elif node.op == "Add" and (isinstance(node.inputs[0], Constant) or isinstance(node.inputs[1], Constant)):
                bias_ind = 0 if isinstance(node.inputs[0], Constant) else 1
                bias, x = node.inputs[bias_ind], node.inputs[1 - bias_ind]
                bias_scale = qparams_per_value[x.name].scale
                qvalues_dict[bias.name] = Constant(bias.name, [], quantize_tensor(bias.data, 4 * bit_width, bias_scale, None))
                qparams_per_value[bias.name] = QuantizationParams(bias_scale, None)
                qparams_per_value[out_val.name] = get_quantization_params(out_val, True)
            elif node.op in ("Identity", "Relu"):
                qparams_per_value[out_val.name] = qparams_per_value[node.inputs[0].name]
            else:
                qparams_per_value[out_val.name] = get_quantization_params(out_val, True)
            qvalues_dict[out_val.name] = Variable(out_val.name, [], [], None)
            qnodes_dict[node.name] = Node(node.name, node.op, node.attrs, [], [])
        for name, qnode in qnodes_dict.items():
            qnode.inputs = [qvalues_dict[i.name] for i in node_dict[name].inputs]
            qnode.outputs = [qvalues_dict[o.name] for o in node_dict[name].outputs]
        for name, qvalue in qvalues_dict.items():
            if isinstance(qvalue, Variable):
                qvalue.inputs = [qnodes_dict[i.name] for i in value_dict[name].inputs]
            qvalue.outputs = [qnodes_dict[o.name] for o in value_dict[name].outputs]
        qoutputs = [qvalues_dict[o.name] for o in self.outputs]
        qinputs = [qvalues_dict[i.name] for i in self.inputs]
        return QModel(list(qnodes_dict.values()), list(qvalues_dict.values()), qinputs, qoutputs, bit_width,
                      qparams_per_value)


class QuantizationParams:
    def __init__(self, scale: np.float32, zero_point):
        self.scale = scale
        self.zero_point = zero_point

    def __repr__(self):
        return f"QuantizationParams(scale={self.scale}, zero_point={self.zero_point})"


class QModel(Model):
    def __init__(self, nodes, values, inputs, outputs, bit_width: int, quant_params: dict):
        """quant_params: value name -> quantization parameter"""
        super().__init__(nodes, values, inputs, outputs)
        self.bit_width = bit_width
        self.quant_params = quant_params
        self._const_deq: dict = {}
        self._plan = None
        # Softmax inside the attention-score GEMM epilogue adds the row sum in a different order than the
        # standalone kernel (float32 addition is not associative): same 1e-5 float contract, but a
        # probability within ~1e-7 of a rounding boundary may quantize to the neighbouring code.  Set to
        # False to keep retain=False bit-identical to the node-by-node run.
        self.fuse_softmax_epilogue = True
        self.fuse_gelu_epilogue = True
        self.fuse_layernorm_glue = True
        self.fuse_attention = True
        self._pipes = {}
        self._graphs: dict = {}
        self._graph_launches: dict = {}

    def __repr__(self):
        return (f"QModel(nodes={self.nodes}, values={self.values}, inputs={self.inputs}, outputs={self.values}, "
                f"bit_width={self.bit_width}, quant_params={self.quant_params})")

    def __str__(self):
        res = "QModel(\n"
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            if isinstance(v, list):
                res += f"  {k}=[\n"
                for e in v:
                    res += f"    {e}\n"
                res += "  ],\n"
            elif isinstance(v, dict):
                res += f"  {k}={{\n"
                for ek, ev in v.items():
                    res += f"    {ek}: {ev},\n"
                res += "  }},\n"
            else:
                res += f"  {k}={v},\n"
        res += ")\n"
        return res

    # ------------------------------------------------------------------ fusion plan (retain=False)
    def _feeds_only_matmul_lhs(self, value: Value) -> bool:
        """True when every consumer of `value` is a MatMul taking it as its left operand, so the
        float tensor itself is never needed -- only its quantized GEMM operand."""
        if not value.outputs or any(value is o for o in self.outputs):
            return False
        return all(n.op == "MatMul" and n.inputs[0] is value and n.inputs[1] is not value for n in value.outputs)

    def _build_plan(self) -> dict:
        """Static multi-node patterns of the graph, looked up by node name at run time:
          gelu[div]      Div -> Erf -> Add -> Mul -> Mul chain        -> one kernel at the Div node
          softmax[div]   Div(x, scalar) -> Softmax(last axis)          -> one kernel at the Div node
          skip           nodes whose work is done by a fused step
          emit[node]     fused result is handed to this node's output
          quantize_out   producers (LayerNorm / Softmax / GELU) whose output only feeds MatMul left
                         operands: they emit the int8 GEMM operand directly instead of float32
        """
        producers = {o.name: n for n in self.nodes for o in n.outputs}
        by_name = {n.name: n for n in self.nodes}
        plan = dict(gelu=find_gelu_chains(self.nodes), softmax={}, skip=set(), emit={}, quantize_out=set(), node=by_name,
                    residual={}, to_operand={}, merge_heads={}, gelu_in={}, attention={}, attention_pv={}, attention_tr={})
        for first, spec in plan["gelu"].items():
            plan["skip"].update(spec[4])
            plan["emit"][spec[5]] = first
        for n in self.nodes:
            if n.op == "Div" and len(n.outputs[0].outputs) == 1 and first_not_in(n.name, plan["gelu"]):
                c = _scalar_const(n.inputs[1], producers)
                sm = n.outputs[0].outputs[0]
                if c is not None and sm.op == "Softmax" and sm.attrs.get("axis", -1) == -1 \
                        and not any(n.outputs[0] is o for o in self.outputs):
                    plan["softmax"][n.name] = (n.inputs[0], c, sm.name)
                    plan["emit"][sm.name] = n.name
        # bias Add (folded into the GEMM epilogue) whose only consumer is a residual Add
        for n in self.nodes:
            if n.op != "Add" or len(n.outputs[0].outputs) != 1 or any(n.outputs[0] is o for o in self.outputs):
                continue
            a, b = n.inputs
            if not ((isinstance(a, Constant) and isinstance(b, Variable)) or (isinstance(b, Constant) and isinstance(a, Variable))):
                continue
            acc_v = b if isinstance(a, Constant) else a
            if not acc_v.inputs or acc_v.inputs[0].op != "MatMul":
                continue
            add2 = n.outputs[0].outputs[0]
            if add2.op != "Add" or add2.name in plan["emit"] or add2.name in plan["skip"]:
                continue
            other = [v for v in add2.inputs if v is not n.outputs[0]]
            if len(other) == 1 and isinstance(other[0], Variable):
                plan["residual"][n.name] = (other[0], add2.name)
        # bias Add -> Reshape([B,S,H,D]) -> Transpose -> MatMul operand: the GEMM epilogue writes the
        # consumer's int8 operand directly (attention Q / K^T / V)
        def single(v):
            return len(v.outputs) == 1 and not any(v is o for o in self.outputs)
        for n in self.nodes:
            if n.op != "Add" or n.name in plan["residual"] or not single(n.outputs[0]):
                continue
            a, b = n.inputs
            if not ((isinstance(a, Constant) and isinstance(b, Variable)) or (isinstance(b, Constant) and isinstance(a, Variable))):
                continue
            rs = n.outputs[0].outputs[0]
            if rs.op != "Reshape" or rs.inputs[0] is not n.outputs[0] or not single(rs.outputs[0]):
                continue
            tr = rs.outputs[0].outputs[0]
            if tr.op != "Transpose" or not single(tr.outputs[0]):
                continue
            mm = tr.outputs[0].outputs[0]
            if mm.op != "MatMul" or mm.inputs[0] is mm.inputs[1]:
                continue
            pos = 0 if mm.inputs[0] is tr.outputs[0] else 1
            perm = [int(x) for x in tr.attrs["perm"]]
            kind = {(0, (0, 2, 1, 3)): "split_rows", (1, (0, 2, 3, 1)): "split_rows", (1, (0, 2, 1, 3)): "split_cols"}.get(
                (pos, tuple(perm)))
            if kind:
                plan["to_operand"][n.name] = dict(reshape=rs, transpose=tr, matmul=mm, pos=pos, kind=kind)
        # P.V MatMul -> Transpose(0,2,1,3) -> Reshape([B,S,H*D]) -> MatMul left operand (attention output proj.)
        for n in self.nodes:
            if n.op != "Transpose" or [int(x) for x in n.attrs["perm"]] != [0, 2, 1, 3] or not single(n.outputs[0]):
                continue
            rs = n.outputs[0].outputs[0]
            if rs.op != "Reshape" or rs.inputs[0] is not n.outputs[0]:
                continue
            if self._feeds_only_matmul_lhs(rs.outputs[0]):
                plan["merge_heads"][n.name] = dict(reshape=rs)
        # Div -> Softmax -> MatMul(P, V) -> Transpose(0,2,1,3) -> Reshape -> MatMul left operand: one attention kernel
        for dname, (xv, c, sm_name) in plan["softmax"].items():
            smv = by_name[sm_name].outputs[0]
            if not single(smv):
                continue
            pv = smv.outputs[0]
            if pv.op != "MatMul" or pv.inputs[0] is not smv or pv.inputs[1] is smv or not single(pv.outputs[0]):
                continue
            tr = pv.outputs[0].outputs[0]
            if tr.name in plan["merge_heads"] and tr.inputs[0] is pv.outputs[0]:
                plan["attention"][dname] = dict(softmax=sm_name, pv=pv.name, transpose=tr.name,
                                                reshape=plan["merge_heads"][tr.name]["reshape"].name)
                plan["attention_pv"][pv.name] = dname
                plan["attention_tr"][tr.name] = dname
        emitters = {}
        for n in self.nodes:
            if n.op == "LayerNormalization" or (n.op == "Softmax" and n.attrs.get("axis", -1) == -1):
                emitters[n.name] = n.outputs[0]
        for first, spec in plan["gelu"].items():
            emitters[spec[5]] = by_name[spec[5]].outputs[0]
        for name, out in emitters.items():
            if self._feeds_only_matmul_lhs(out):
                plan["quantize_out"].add(name)
        # bias Add whose value is consumed only by a GELU chain that feeds MatMul left operands: the GEMM
        # producing the Add's accumulator can run bias + GELU + quantize in its epilogue
        for first, spec in plan["gelu"].items():
            x, last = spec[0], spec[5]
            if last not in plan["quantize_out"] or not isinstance(x, Variable) or not x.inputs:
                continue
            add = x.inputs[0]
            chain = {first, *spec[4], last}
            if add.op == "Add" and all(c.name in chain for c in x.outputs) and not any(x is o for o in self.outputs) \
                    and add.name not in plan["residual"] and add.name not in plan["to_operand"]:
                plan["gelu_in"][add.name] = first
        return plan

    # ------------------------------------------------------------------------------
    def _dequantized(self, value: Value) -> FTensor:
        """Dequantize a value for a float consumer; constants are immutable, so cached."""
        if isinstance(value, Constant):
            d = self._const_deq.get(value.name)
            if d is None or d[0] is not value.data:
                d = (value.data, value.data.dequantize())
                self._const_deq[value.name] = d
            return d[1]
        return value.data.dequantize()

    def _row_gather_after(self, node: Node):
        """(Gather node, index) if the only consumer of this LayerNormalization (last axis, 3-D input) is a Gather along
        axis 1 with one constant scalar index and the normalised value is not a model output, else None."""
        out = node.outputs[0]
        x = node.inputs[0].data
        if out in self.outputs or len(out.outputs) != 1 or not isinstance(x, FTensor) or x.device_tensor.dim() != 3 \
                or node.attrs.get("axis", -1) not in (-1, 2):
            return None
        g = out.outputs[0]
        if g.op != "Gather" or g.inputs[0] is not out or int(g.attrs.get("axis", 0)) != 1:
            return None
        idx = g.inputs[1].data
        if not isinstance(idx, ITensor) or np.asarray(idx.data).ndim != 0:
            return None
        i = int(np.asarray(idx.data))
        n = int(x.device_tensor.shape[1])
        if i < 0:
            i += n
        return (g, i) if 0 <= i < n else None

    def _conv_input_tiling(self, value: Value):
        """(kh, kw) if every Conv consuming `value` (see _conv_input_pads) has kernel == strides == (kh, kw), else None."""
        kernel = None
        for n in value.outputs:
            w = n.inputs[1].data
            k = tuple(int(v) for v in w.shape[2:])
            if len(k) != 2 or tuple(int(v) for v in n.attrs["strides"]) != k or (kernel is not None and k != kernel):
                return None
            kernel = k
        return kernel

    def _conv_input_pads(self, value: Value):
        """The common `pads` of the Conv nodes consuming `value` as their image, or None if anything else reads it
        (model outputs included) or the quantized weights are not symmetric 8-bit-or-narrower codes."""
        if not value.outputs or value in self.outputs:
            return None
        pads = None
        for n in value.outputs:
            if n.op != "Conv" or n.inputs[0] is not value or value in n.inputs[1:]:
                return None
            w = n.inputs[1].data
            if not isinstance(w, QTensor) or w._zp is not None or w.bit_width > 8:
                return None
            p = tuple(int(v) for v in n.attrs["pads"])
            if pads is not None and p != pads:
                return None
            pads = p
        return pads

    def _rowsum_needed(self, value: Value) -> bool:
        """Row sums of a left operand are needed iff some consuming MatMul has an asymmetric right operand."""
        for n in value.outputs:
            if n.op in ("MatMul", "Gemm") and n.inputs[0] is value:
                other = n.inputs[1]
                if isinstance(other.data, QTensor):
                    if other.data._zp is not None:
                        return True
                elif self.quant_params[other.name].zero_point is not None:
                    return True
        return False

    def _quantized_operand(self, value: Value, role: Optional[str], other: Value, cache: dict) -> QTensor:
        """Quantize a float activation for an integer MatMul / Gemm (model.py:503-527), straight
        into the GEMM operand layout; one quantization per (value, role) and forward."""
        key = (value.name, role)
        q = cache.get(key)
        if q is None:
            qp = self.quant_params[value.name]
            if isinstance(other.data, QTensor):
                other_asym = other.data._zp is not None
            else:
                other_asym = self.quant_params[other.name].zero_point is not None
            q = quantize_tensor(value.data, self.bit_width, qp.scale, qp.zero_point, role=role, want_rowsum=other_asym)
            cache[key] = q
        return q

    def _emit_operand(self, out: Value, op, shape, cache: dict) -> None:
        """Register a fused producer's int8 operand as the quantized form of `out` (left operand)."""
        qp = self.quant_params[out.name]
        cache[(out.name, "A")] = qtensor_from_operand(op, "A", shape, self.bit_width, qp.scale, qp.zero_point)

    # ------------------------------------------------------------------ CUDA-graph replay
    def _graph_call(self, inputs: list, device_outputs: bool):
        """Capture the fused forward (retain=False) once per input signature into a CUDA graph and
        replay it: ~200 kernel launches become one graph launch, so the host interpreter loop
        (Python + ctypes per node) disappears from the steady state.  Quantization parameters are
        static after calibration, so the launch sequence depends on shapes only."""
        key = tuple((tuple(a.shape), str(a.dtype)) for a in inputs) + (self.fuse_softmax_epilogue, self.fuse_gelu_epilogue, self.fuse_layernorm_glue, self.fuse_attention)
        entry = self._graphs.get(key)
        dev = torch.device("cuda", torch.cuda.current_device())
        if entry is None:
            for a in inputs:
                if not (isinstance(a, torch.Tensor) or a.dtype == np.float32):
                    raise ValueError("graph=True supports float32 inputs only")
            static_in = [torch.empty(tuple(a.shape), dtype=torch.float32, device=dev) for a in inputs]
            for s_in, a in zip(static_in, inputs):
                s_in.copy_(a if isinstance(a, torch.Tensor) else torch.from_numpy(a))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                     # warm-up outside capture: constants, operand
                for _ in range(2):                            # caches, function attributes, allocator pool
                    self(static_in, retain=False, device_outputs=True)
            torch.cuda.current_stream().wait_stream(side)
            self.release()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self(static_in, retain=False, device_outputs=True)
            self.release()
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        for s_in, a in zip(static_in, inputs):
            if isinstance(a, torch.Tensor):
                if a.data_ptr() != s_in.data_ptr():
                    s_in.copy_(a, non_blocking=True)
            else:
                s_in.copy_(torch.from_numpy(a), non_blocking=True)
        graph.replay()
        K._count(self._graph_launches.setdefault(key, 0))
        if device_outputs:
            return list(static_out)
        return [o.cpu().numpy() for o in static_out]

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """Write the quantized model -- graph, quantization parameters and the quantized constants with their
        bit_width-bit codes packed (nq_pack_s8) -- as one `.npz`; `QModel.load` restores it without the float
        weights or a calibration pass.  (The reference has no persistence of `QModel`.)"""
        import json
        arrays: dict = {}

        def enc_attr(v):
            if isinstance(v, np.ndarray):
                key = f"attr{len(arrays)}"
                arrays[key] = v
                return {"t": "array", "k": key}
            if isinstance(v, (bytes, bytearray)):
                return {"t": "bytes", "v": v.decode("latin1")}
            if isinstance(v, (list, tuple)):
                return {"t": "list", "v": [enc_attr(x) for x in v]}
            if isinstance(v, (np.integer, int)):
                return {"t": "int", "v": int(v)}
            if isinstance(v, (np.floating, float)):
                return {"t": "float", "v": float(v)}
            if isinstance(v, str):
                return {"t": "str", "v": v}
            raise ValueError(f"cannot serialise attribute of type {type(v)}")

        def enc_zp(z):
            return None if z is None else int(np.asarray(z).reshape(-1)[0])

        values = []
        for i, v in enumerate(self.values):
            rec = {"name": v.name, "kind": "const" if isinstance(v, Constant) else "var"}
            d = v.data if isinstance(v, Constant) else None
            if isinstance(d, QTensor):
                codes = np.ascontiguousarray(d.data)                       # int64, reference layout
                rec.update(tensor="q", bits=int(d.bit_width), scale=float(np.float32(d.scale)), zp=enc_zp(d.zero_point),
                           shape=list(codes.shape))
                key = f"const{i}"
                if d.bit_width <= 8 and codes.size:
                    packed = K.pack(_to_device(codes.astype(np.int8)).reshape(-1), int(d.bit_width))
                    arrays[key] = packed.cpu().numpy()
                    rec["packed"] = True
                else:
                    arrays[key] = codes
                    rec["packed"] = False
                rec["k"] = key
            elif isinstance(d, ITensor):
                key = f"const{i}"
                arrays[key] = np.asarray(d.data)
                rec.update(tensor="i", k=key)
            elif isinstance(d, FTensor):
                key = f"const{i}"
                arrays[key] = np.asarray(d.data)
                rec.update(tensor="f", k=key)
            values.append(rec)
        meta = {
            "format": "numpy_quant_b200.qmodel/1", "bit_width": int(self.bit_width),
            "nodes": [{"name": n.name, "op": n.op, "attrs": {k: enc_attr(a) for k, a in n.attrs.items()},
                       "inputs": [i.name for i in n.inputs], "outputs": [o.name for o in n.outputs]} for n in self.nodes],
            "values": values, "inputs": [v.name for v in self.inputs], "outputs": [v.name for v in self.outputs],
            "qparams": {name: [float(np.float32(qp.scale)), enc_zp(qp.zero_point)] for name, qp in self.quant_params.items()},
        }
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez(path, **arrays)

    @classmethod
    def load(cls, path: str) -> "QModel":
        """Restore a model written by `save` (codes unpacked on the device)."""
        import json
        with np.load(path if str(path).endswith(".npz") else str(path) + ".npz") as z:
            meta = json.loads(bytes(z["meta"]).decode())
            if meta.get("format") != "numpy_quant_b200.qmodel/1":
                raise ValueError("not a numpy_quant_b200 quantized-model file")
            arrays = {k: z[k] for k in z.files if k != "meta"}

        def dec_attr(e):
            t = e["t"]
            if t == "array":
                return arrays[e["k"]]
            if t == "bytes":
                return e["v"].encode("latin1")
            if t == "list":
                return [dec_attr(x) for x in e["v"]]
            return e["v"]

        def dec_zp(zv):
            return None if zv is None else np.int64(zv)

        vals: dict = {}
        for rec in meta["values"]:
            if rec["kind"] == "var":
                vals[rec["name"]] = Variable(rec["name"], [], [], None)
                continue
            kind = rec.get("tensor")
            if kind == "q":
                shape = tuple(rec["shape"])
                if rec["packed"]:
                    n = int(np.prod(shape))
                    codes = K.unpack(_to_device(arrays[rec["k"]]), n, int(rec["bits"])).view(shape)
                    data = QTensor(codes, int(rec["bits"]), np.float32(rec["scale"]), dec_zp(rec["zp"]))
                else:
                    data = QTensor(arrays[rec["k"]].astype(np.int64).reshape(shape), int(rec["bits"]), np.float32(rec["scale"]),
                                   dec_zp(rec["zp"]))
            elif kind == "i":
                data = ITensor(arrays[rec["k"]])
            elif kind == "f":
                data = FTensor(arrays[rec["k"]])
            else:
                data = None
            vals[rec["name"]] = Constant(rec["name"], [], data)
        nodes = []
        for nrec in meta["nodes"]:
            node = Node(nrec["name"], nrec["op"], {k: dec_attr(a) for k, a in nrec["attrs"].items()},
                        [vals[i] for i in nrec["inputs"]], [vals[o] for o in nrec["outputs"]])
            nodes.append(node)
            for v in node.inputs:
                v.outputs.append(node)
            for v in node.outputs:
                if isinstance(v, Variable):
                    v.inputs.append(node)
        qparams = {name: QuantizationParams(np.float32(sv), dec_zp(zv)) for name, (sv, zv) in meta["qparams"].items()}
        return cls(nodes, [vals[r["name"]] for r in meta["values"]], [vals[n] for n in meta["inputs"]],
                   [vals[n] for n in meta["outputs"]], int(meta["bit_width"]), qparams)

    # ------------------------------------------------------------------ sub-byte weight storage
    def pack_weights(self) -> dict:
        """Store every quantized MatMul / Gemm weight (right operand, a Constant) as a bit_width-bit packed
        bitstream (BASELINE config 4: int4 / 2-bit sweeps with sub-byte packing); each forward unpacks a weight
        into a transient int8 operand right before its GEMM.  Results are unchanged.  Returns the byte counts."""
        packed = dense = 0
        for n in self.nodes:
            if n.op not in ("MatMul", "Gemm") or len(n.inputs) < 2:
                continue
            wv = n.inputs[1]
            if not isinstance(wv, Constant) or not isinstance(wv.data, QTensor) or wv.data.bit_width > 8:
                continue
            # Gemm(transB=1) consumes the transposed view of the stored [N, K] matrix: its K-major form is role "A"
            role = "A" if (n.op == "Gemm" and n.attrs.get("transB")) else "B"
            q = wv.data
            if len(q.shape) != 2:
                continue
            dense += int(np.prod(q.shape))
            packed += q.pack_storage(role)
        self._graphs.clear()                                 # captured graphs hold the old operand buffers
        self._pipes.clear()
        return {"int8_bytes": dense, "resident_bytes": packed, "bit_width": self.bit_width}

    # ------------------------------------------------------------------ pipelined host <-> device serving
    def submit(self, inputs: list) -> "PendingOutputs":
        """Asynchronous graph-replay forward for host inputs: returns at once with a handle whose `.result()`
        yields the host outputs.  Host->device copies run on their own stream into one of two staging buffers,
        device->host copies on a third stream into pinned buffers, so with two submissions in flight the
        copies of step k+1 / k-1 overlap the kernels of step k.  Same results as `self(inputs, graph=True)`."""
        key = tuple((tuple(a.shape), str(a.dtype)) for a in inputs) + (self.fuse_softmax_epilogue, self.fuse_gelu_epilogue, self.fuse_layernorm_glue, self.fuse_attention)
        if key not in self._graphs:
            outs = self(inputs, graph=True)                   # first use: warm-up + capture, synchronous
            return PendingOutputs(outs, None)
        graph, static_in, static_out = self._graphs[key]
        pipe = self._pipes.get(key)
        if pipe is None:
            pipe = dict(h2d=torch.cuda.Stream(), d2h=torch.cuda.Stream(), n=0,
                        stage=[[torch.empty_like(t) for t in static_in] for _ in range(2)],
                        host=[[torch.empty(tuple(o.shape), dtype=o.dtype, pin_memory=True) for o in static_out] for _ in range(2)],
                        consumed=[None, None], done=[None, None])
            self._pipes[key] = pipe
        slot = pipe["n"] % 2
        pipe["n"] += 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(pipe["h2d"]):
            if pipe["consumed"][slot] is not None:
                pipe["h2d"].wait_event(pipe["consumed"][slot])      # staging buffer free again
            for dst, a in zip(pipe["stage"][slot], inputs):
                dst.copy_(a if isinstance(a, torch.Tensor) else torch.from_numpy(a), non_blocking=True)
            staged = torch.cuda.Event()
            staged.record(pipe["h2d"])
        main.wait_event(staged)
        for s_in, st in zip(static_in, pipe["stage"][slot]):
            s_in.copy_(st, non_blocking=True)
        consumed = torch.cuda.Event()
        consumed.record(main)
        pipe["consumed"][slot] = consumed
        other = pipe["done"][1 - slot]
        if other is not None:
            main.wait_event(other)                                  # previous outputs have left static_out
        graph.replay()
        K._count(self._graph_launches.setdefault(key, 0))
        computed = torch.cuda.Event()
        computed.record(main)
        with torch.cuda.stream(pipe["d2h"]):
            pipe["d2h"].wait_event(computed)
            for h, o in zip(pipe["host"][slot], static_out):
                h.copy_(o, non_blocking=True)
            done = torch.cuda.Event()
            done.record(pipe["d2h"])
        pipe["done"][slot] = done
        return PendingOutputs(pipe["host"][slot], done)

    def __call__(self, inputs: List[np.ndarray], profile=False, *, retain: bool = True,
                 device_outputs: bool = False, graph: bool = False):
        """Quantized interpreter (reference model.py:486-565).

        retain=True  keeps every intermediate on `Variable.data` like the reference.
        retain=False frees a value after its last consumer and fuses multi-node patterns
                     (GELU chain, Div+Softmax, producer->quantize); results are bit-identical.
        device_outputs=True returns CUDA tensors (no device->host copy, no synchronisation).
        graph=True   (implies retain=False) replays the forward as one CUDA graph.
        """
        if graph:
            if profile:
                raise ValueError("profile=True needs the eager interpreter (graph=False)")
            if not torch.cuda.is_current_stream_capturing():
                key = tuple((tuple(a.shape), str(a.dtype)) for a in inputs) + (self.fuse_softmax_epilogue, self.fuse_gelu_epilogue, self.fuse_layernorm_glue, self.fuse_attention)
                if key not in self._graphs:
                    before = K.LAUNCHES
                    out = self._graph_call(inputs, device_outputs)
                    # launches of one forward = (2 warm-ups + 1 capture) / 3
                    self._graph_launches[key] = (K.LAUNCHES - before) // 3
                    return out
                return self._graph_call(inputs, device_outputs)
        for array, variable in zip(inputs, self.inputs):
            qparams = self.quant_params[variable.name]
            if isinstance(array, torch.Tensor) or array.dtype == np.float32:
                pads = self._conv_input_pads(variable)
                q = None
                if pads is not None:
                    tiling = self._conv_input_tiling(variable) if not any(pads) else None
                    if tiling is not None:
                        # kernel == stride, no padding (patch embedding): quantized straight into the patch matrix
                        q = quantize_tensor_patches(FTensor(array), self.bit_width, qparams.scale, qparams.zero_point, tiling)
                    if q is None:
                        # consumed only by Conv nodes: quantized straight into the padded NHWC image their TMA reads
                        q = quantize_tensor_nhwc(FTensor(array), self.bit_width, qparams.scale, qparams.zero_point, pads)
                variable.data = q if q is not None else \
                    quantize_tensor(FTensor(array), self.bit_width, qparams.scale, qparams.zero_point)
            elif array.dtype == np.int64:
                variable.data = ITensor(array)
            else:
                raise ValueError(f"Array dtype {array.dtype} not supported")

        times = {op: 0.0 for op in {n.op for n in self.nodes}}
        times["TinyqQuant"] = 0.0
        times["TinyqDequant"] = 0.0

        def tick():
            if profile:
                torch.cuda.synchronize()
                return time()
            return 0.0

        def tock(key, t0):
            if profile:
                torch.cuda.synchronize()
                times[key] += time() - t0

        fused = not retain
        if fused and self._plan is None:
            self._plan = self._build_plan()
        plan = self._plan if fused else dict(gelu={}, softmax={}, skip=set(), emit={}, quantize_out=set(), node={},
                                             residual={}, to_operand={}, merge_heads={}, gelu_in={}, attention={}, attention_pv={}, attention_tr={})
        dyn_skip: set = set()
        attn_pending: dict = {}
        qcache: dict = {}
        stash: dict = {}
        remaining = None
        if fused:
            remaining = {v.name: len(v.outputs) for v in self.values if isinstance(v, Variable)}
            keep = {o.name for o in self.outputs} | {n.outputs[0].name for n in self.nodes if n.op == "Constant"}
        bits = self.bit_width

        def as_float(v: Value) -> FTensor:
            return v.data if isinstance(v.data, FTensor) else self._dequantized(v)

        for node in self.nodes:
            name = node.name
            out0 = node.outputs[0] if node.outputs else None
            if name in dyn_skip:
                outputs_data = [None]
            elif name in plan["gelu"]:
                # ---- GELU chain in one kernel (optionally emitting the next MatMul's int8 operand)
                x, c1, c2, c3, _, last = plan["gelu"][name]
                xin = as_float(x)
                t0 = tick()
                lastv = plan["node"][last].outputs[0] if last in plan["quantize_out"] else None
                if lastv is not None and xin.device_tensor.dim() >= 2:
                    qp = self.quant_params[lastv.name]
                    op = K.gelu_quantize(xin.device_tensor, c1, c2, c3, bits, float(qp.scale), _zp_int(qp.zero_point),
                                         self._rowsum_needed(lastv))
                    self._emit_operand(lastv, op, tuple(xin.device_tensor.shape), qcache)
                    stash[last] = None
                else:
                    stash[last] = xin.gelu_erf(c1, c2, c3)
                tock("Erf", t0)
                outputs_data = [None]
            elif name in plan["softmax"]:
                # ---- Div + Softmax (+ quantize) in one kernel
                x, c, sm_name = plan["softmax"][name]
                smv = plan["node"][sm_name].outputs[0]
                aspec = plan["attention"].get(name) if (self.fuse_attention and self.fuse_softmax_epilogue) else None
                if aspec is not None and isinstance(x.data, QTensor) and x.data._pending():
                    # whole attention in one kernel: remember the pending score GEMM, launch at the merge-heads node
                    L = x.data._lazy
                    if L.get("bias_q") is None and L["a"].batch == int(np.prod(L["batch_shape"] or (1,))) \
                            and K.can_fuse_attention_qk(L["a"], L["b"], getattr(x.data._zp, "zp_a", None), getattr(x.data._zp, "zp_b", None)):
                        attn_pending[name] = dict(scores=x.data, c=c, spec=aspec)
                        stash[sm_name] = None
                        for o in node.outputs:
                            o.data = None
                        self._release_inputs(node, remaining, keep, qcache)
                        continue
                if self.fuse_softmax_epilogue and sm_name in plan["quantize_out"] and isinstance(x.data, QTensor) \
                        and x.data._pending():
                    # scores never leave the GEMM: softmax + quantize run in its epilogue
                    t0 = tick()
                    qp = self.quant_params[smv.name]
                    qP = x.data.softmax_into_operand(c, bits, qp.scale, qp.zero_point, self._rowsum_needed(smv))
                    tock("Softmax", t0)
                    if qP is not None:
                        qcache[(smv.name, "A")] = qP
                        stash[sm_name] = None
                        for o, tensor in zip(node.outputs, [None]):
                            o.data = tensor
                        self._release_inputs(node, remaining, keep, qcache)
                        continue
                xin = as_float(x)
                t0 = tick()
                xt = xin.device_tensor
                if sm_name in plan["quantize_out"] and K.can_fuse_softmax_quantize(xt):
                    qp = self.quant_params[smv.name]
                    op = K.softmax_quantize(xt, c, bits, float(qp.scale), _zp_int(qp.zero_point), self._rowsum_needed(smv))
                    self._emit_operand(smv, op, tuple(xt.shape), qcache)
                    stash[sm_name] = None
                else:
                    stash[sm_name] = FTensor(K.softmax_div_lastdim(xt, c))
                tock("Softmax", t0)
                outputs_data = [None]
            elif name in plan["skip"] or name in dyn_skip:
                outputs_data = [None]
            elif name in plan["emit"] or name in stash:
                outputs_data = [stash.pop(name)]
            elif name in plan["attention_pv"] and plan["attention_pv"][name] in attn_pending:
                # P.V of a fused attention: only collect the V operand (already in its K-major layout)
                rec = attn_pending[plan["attention_pv"][name]]
                vval = node.inputs[1]
                vq = qcache.get((vval.name, "B"))
                if vq is None or isinstance(vval.data, QTensor):
                    if isinstance(vval.data, FTensor):
                        t0 = tick()
                        vq = self._quantized_operand(vval, "B", node.inputs[0], qcache)
                        tock("TinyqQuant", t0)
                    else:
                        vq = vval.data
                rec["v"] = vq
                outputs_data = [None]
            elif name in plan["attention_tr"] and plan["attention_tr"][name] in attn_pending:
                # Transpose(0,2,1,3) of the context: launch the attention kernel, emit the output projection's operand
                rec = attn_pending.pop(plan["attention_tr"][name])
                spec = rec["spec"]
                smv = plan["node"][spec["softmax"]].outputs[0]
                outv = plan["node"][spec["reshape"]].outputs[0]
                qp_p, qp_o = self.quant_params[smv.name], self.quant_params[outv.name]
                t0 = tick()
                try:
                    qctx = rec["scores"].attention_into_operand(rec["c"], rec["v"], bits, qp_p.scale, qp_p.zero_point, qp_o.scale,
                                                                qp_o.zero_point, self._rowsum_needed(outv))
                except (_lib.NqError, ValueError):
                    # parameters outside the fused kernel's host-checked windows (e.g. a softmax range calibrated far
                    # from zero, huge zero-points): the pending score GEMM is still alive -- two-GEMM route
                    qctx = self._attention_unfused(rec, smv, outv, qp_p, qp_o, bits)
                tock("MatMul", t0)
                if isinstance(qctx, QTensor):
                    qcache[(outv.name, "A")] = qctx
                    stash[spec["reshape"]] = None
                else:
                    stash[spec["reshape"]] = qctx                # float32 context [B, S, H*D]: quantized by its consumer
                outputs_data = [None]
            elif node.op == "Constant" and out0.data is not None:
                outputs_data = [out0.data]                      # immutable: uploaded once, reused
            elif node.op in ("MatMul", "Gemm"):
                inputs_data = []
                for pos, i in enumerate(node.inputs):
                    role = None
                    if pos < 2:
                        trans = node.op == "Gemm" and node.attrs.get("transA" if pos == 0 else "transB")
                        role = ("A" if pos == 0 else "B") if not trans else None
                    pre = qcache.get((i.name, role)) if role else None
                    if pre is not None and not isinstance(i.data, QTensor):
                        inputs_data.append(pre)
                    elif isinstance(i.data, FTensor):
                        t0 = tick()
                        if role is not None and i.data.device_tensor.dim() >= 2:
                            inputs_data.append(self._quantized_operand(i, role, node.inputs[1 - pos], qcache))
                        else:
                            qp = self.quant_params[i.name]
                            inputs_data.append(quantize_tensor(i.data, bits, qp.scale, qp.zero_point))
                        tock("TinyqQuant", t0)
                    else:
                        inputs_data.append(i.data)
                t0 = tick()
                outputs_data = onnx_operator_implementation(node.op, inputs_data, node.attrs)
                if node.op == "Gemm":
                    qp = self.quant_params[out0.name]
                    outputs_data = [outputs_data[0].requantize(bits, qp.scale, qp.zero_point)]
                tock(node.op, t0)
            elif node.op == "Add" and self._is_bias_add(node):
                # dequantize(acc) + dequantize(bias): both folded into the GEMM epilogue
                acc, bias = (node.inputs[0], node.inputs[1]) if isinstance(node.inputs[1], Constant) else \
                            (node.inputs[1], node.inputs[0])
                t0 = tick()
                b = self._dequantized(bias)
                tock("TinyqDequant", t0)
                t0 = tick()
                spec = plan["to_operand"].get(name)
                if spec is not None and self._emit_split_heads(acc.data, b, spec, qcache):
                    dyn_skip.update((spec["reshape"].name, spec["transpose"].name))
                    outputs_data = [None]
                    tock(node.op, t0)
                    for o, tensor in zip(node.outputs, outputs_data):
                        o.data = tensor
                    self._release_inputs(node, remaining, keep, qcache)
                    continue
                gspec = plan["gelu_in"].get(name) if self.fuse_gelu_epilogue else None
                if gspec is not None:
                    # bias Add -> GELU chain -> MatMul left operand: all in this GEMM's epilogue
                    first, (_, c1, c2, c3, _, last) = gspec, plan["gelu"][gspec]
                    lastv = plan["node"][last].outputs[0]
                    qp = self.quant_params[lastv.name]
                    qg = acc.data.gelu_into_operand(b, (c1, c2, c3), bits, qp.scale, qp.zero_point, self._rowsum_needed(lastv))
                    if qg is not None:
                        qcache[(lastv.name, "A")] = qg
                        stash[last] = None
                        dyn_skip.add(first)
                        outputs_data = [None]
                        tock(node.op, t0)
                        for o, tensor in zip(node.outputs, outputs_data):
                            o.data = tensor
                        self._release_inputs(node, remaining, keep, qcache)
                        continue
                res_spec = plan["residual"].get(name)
                resid = res_spec[0].data if res_spec else None
                if isinstance(resid, FTensor) and tuple(resid.device_tensor.shape) == tuple(acc.data.shape):
                    # (bias + dequant) + residual in the same epilogue; handed to the residual Add's output
                    stash[res_spec[1]] = acc.data.dequantize(bias=b, residual=resid)
                    outputs_data = [None]
                else:
                    outputs_data = [acc.data.dequantize(bias=b)]
                tock(node.op, t0)
            elif fused and name in plan["merge_heads"] and isinstance(node.inputs[0].data, QTensor) \
                    and node.inputs[0].data._pending() and self._emit_merge_heads(node, plan["merge_heads"][name], qcache, tick, tock):
                dyn_skip.add(plan["merge_heads"][name]["reshape"].name)
                outputs_data = [None]
            elif node.op == "Transpose" and isinstance(node.inputs[0].data, QTensor) and node.inputs[0].data._pending() \
                    and list(node.attrs["perm"]) == [0, 2, 1, 3] \
                    and (heads_last := self._timed_heads_last(node.inputs[0].data, tick, tock)) is not None:
                # attention context: the P.V GEMM writes [B, S, H, D] directly, no transpose copy
                outputs_data = [heads_last]
            elif node.op == "Conv" and isinstance(node.inputs[0].data, QTensor) and isinstance(node.inputs[1].data, QTensor):
                t0 = tick()
                b = self._dequantized(node.inputs[2])
                tock("TinyqDequant", t0)
                t0 = tick()
                outputs_data = onnx_operator_implementation(node.op, [node.inputs[0].data, node.inputs[1].data, b], node.attrs)
                tock(node.op, t0)
            elif fused and node.op == "LayerNormalization" and isinstance(node.inputs[0].data, FTensor) \
                    and (gsel := self._row_gather_after(node)) is not None:
                # LayerNormalization -> Gather(axis=1, one index) (the class token in front of the classifier): the
                # normalisation is row-wise, so gathering first gives the same rows with the same arithmetic and the
                # other S - 1 rows of the last LayerNorm are never produced (nothing else reads them: retain=False)
                gnode, idx = gsel
                t0 = tick()
                g, b = as_float(node.inputs[1]), as_float(node.inputs[2])
                xt = node.inputs[0].data.device_tensor
                sub = FTensor(K.materialize(xt[:, idx:idx + 1, :]))
                y = onnx_operator_implementation(node.op, [sub, g, b], node.attrs)[0]
                stash[gnode.name] = FTensor(y.device_tensor.reshape(xt.shape[0], xt.shape[2]))
                outputs_data = [None]
                tock(node.op, t0)
            elif fused and node.op == "LayerNormalization" and name in plan["quantize_out"] \
                    and isinstance(node.inputs[0].data, FTensor) \
                    and K.can_fuse_layernorm_quantize(node.inputs[0].data.device_tensor) \
                    and node.attrs["axis"] in (-1, node.inputs[0].data.device_tensor.dim() - 1):
                # ---- LayerNorm -> quantize: only the int8 operand is written
                t0 = tick()
                g, b = as_float(node.inputs[1]), as_float(node.inputs[2])
                tock("TinyqDequant", t0)
                t0 = tick()
                xt = node.inputs[0].data.device_tensor
                qp = self.quant_params[out0.name]
                op = K.layernorm_quantize(xt, g.device_tensor, b.device_tensor, node.attrs["epsilon"], bits,
                                          float(qp.scale), _zp_int(qp.zero_point), self._rowsum_needed(out0),
                                          float_glue=self.fuse_layernorm_glue)
                self._emit_operand(out0, op, tuple(xt.shape), qcache)
                outputs_data = [None]
                tock(node.op, t0)
            elif fused and node.op == "Softmax" and name in plan["quantize_out"] and isinstance(node.inputs[0].data, FTensor) \
                    and node.attrs.get("axis", -1) == -1 and K.can_fuse_softmax_quantize(node.inputs[0].data.device_tensor):
                t0 = tick()
                xt = node.inputs[0].data.device_tensor
                qp = self.quant_params[out0.name]
                op = K.softmax_quantize(xt, None, bits, float(qp.scale), _zp_int(qp.zero_point), self._rowsum_needed(out0))
                self._emit_operand(out0, op, tuple(xt.shape), qcache)
                outputs_data = [None]
                tock(node.op, t0)
            else:
                inputs_data = []
                for i in node.inputs:
                    if isinstance(i.data, QTensor):
                        t0 = tick()
                        inputs_data.append(self._dequantized(i))
                        tock("TinyqDequant", t0)
                    else:
                        inputs_data.append(i.data)
                t0 = tick()
                outputs_data = onnx_operator_implementation(node.op, inputs_data, node.attrs)
                tock(node.op, t0)

            for o, tensor in zip(node.outputs, outputs_data):
                o.data = tensor
            if remaining is not None:
                self._release_inputs(node, remaining, keep, qcache)

        output_tensors = []
        for out_var in self.outputs:
            if isinstance(out_var.data, FTensor):
                res = out_var.data
            elif isinstance(out_var.data, QTensor):
                res = out_var.data.dequantize()
            else:
                raise ValueError
            output_tensors.append(res.device_tensor if device_outputs else res.data)
        if profile:
            return output_tensors, times
        return output_tensors

    def _release_inputs(self, node: Node, remaining: dict, keep: set, qcache: dict) -> None:
        """retain=False: drop a value (and its cached operands) once its last consumer has run."""
        for i in node.inputs:
            if i.name in remaining:
                remaining[i.name] -= 1
                if remaining[i.name] <= 0 and i.name not in keep and not any(i is v for v in self.inputs):
                    i.data = None
                    qcache.pop((i.name, "A"), None)
                    qcache.pop((i.name, "B"), None)

    def _attention_unfused(self, rec: dict, smv: Value, outv: Value, qp_p, qp_o, bits: int):
        """The attention block without nq_attention_s8, from the pending score GEMM: softmax in the score GEMM's epilogue
        (or the separate softmax -> quantize kernel), P.V as its own GEMM, merge heads + quantize in its epilogue (or as
        float32 when even that does not apply).  Returns the [B, S, H*D] QTensor operand or an FTensor."""
        scores, c, v = rec["scores"], rec["c"], rec["v"]
        qP = scores.softmax_into_operand(c, bits, qp_p.scale, qp_p.zero_point, True) if self.fuse_softmax_epilogue else None
        if qP is None:
            xt = scores.dequantize().device_tensor
            if K.can_fuse_softmax_quantize(xt):
                op = K.softmax_quantize(xt, c, bits, float(qp_p.scale), _zp_int(qp_p.zero_point), True)
                qP = qtensor_from_operand(op, "A", tuple(xt.shape), bits, qp_p.scale, qp_p.zero_point)
            else:
                qP = quantize_tensor(FTensor(K.softmax_div_lastdim(xt, c)), bits, qp_p.scale, qp_p.zero_point, role="A",
                                     want_rowsum=True)
        acc = qP.matmul(v)
        B, H, S, D = (int(x) for x in acc.shape)
        q = acc.quantize_into_operand(None, bits, qp_o.scale, qp_o.zero_point, "merge_heads", H, S, self._rowsum_needed(outv),
                                      "A", (B, S, H * D))
        if q is not None:
            return q
        ctx = acc.dequantize_heads_last()
        if ctx is None:
            ctx = FTensor(K.materialize(acc.dequantize().device_tensor.permute(0, 2, 1, 3)))
        return FTensor(ctx.device_tensor.reshape(B, S, H * D))

    def _emit_split_heads(self, acc: QTensor, bias: FTensor, spec: dict, qcache: dict) -> bool:
        """bias Add -> Reshape -> Transpose -> MatMul operand, inside the GEMM epilogue."""
        shape = spec["reshape"].inputs[1].data
        if not isinstance(shape, ITensor) or np.asarray(shape.data).size != 4:
            return False
        B, S, H, D = (int(x) for x in np.asarray(shape.data).reshape(-1))
        if min(B, S, H, D) <= 0 or tuple(acc.shape[-1:]) != (H * D,) or int(np.prod(acc.shape[:-1])) != B * S:
            return False
        tr_out = spec["transpose"].outputs[0]
        mm, pos = spec["matmul"], spec["pos"]
        role = "A" if pos == 0 else "B"
        qp = self.quant_params[tr_out.name]
        other = mm.inputs[1 - pos]
        if isinstance(other.data, QTensor):
            other_asym = other.data._zp is not None
        else:
            other_asym = self.quant_params[other.name].zero_point is not None
        perm = [int(x) for x in spec["transpose"].attrs["perm"]]
        logical = tuple([B, S, H, D][p] for p in perm)
        # column sums of V: the fused attention kernel does not read them (its zero-point terms come from constant-
        # operand MMAs) and QTensor._operand computes them on demand for the two-GEMM route
        want_rs = other_asym and spec["kind"] != "split_cols"
        q = acc.quantize_into_operand(bias, self.bit_width, qp.scale, qp.zero_point, spec["kind"], H, S, want_rs,
                                      role, logical)
        if q is None:
            return False
        qcache[(tr_out.name, role)] = q
        return True

    def _emit_merge_heads(self, node: Node, spec: dict, qcache: dict, tick, tock) -> bool:
        """P.V accumulator -> Transpose(0,2,1,3) -> Reshape -> left operand of the output projection."""
        acc = node.inputs[0].data
        if len(acc.shape) != 4:
            return False
        B, H, S, D = (int(x) for x in acc.shape)
        out = spec["reshape"].outputs[0]
        qp = self.quant_params[out.name]
        t0 = tick()
        q = acc.quantize_into_operand(None, self.bit_width, qp.scale, qp.zero_point, "merge_heads", H, S,
                                      self._rowsum_needed(out), "A", (B, S, H * D))
        tock("TinyqQuant", t0)
        if q is None:
            return False
        qcache[(out.name, "A")] = q
        return True

    @staticmethod
    def _timed_heads_last(acc: QTensor, tick, tock):
        t0 = tick()
        res = acc.dequantize_heads_last()
        tock("TinyqDequant", t0)
        return res

    @staticmethod
    def _is_bias_add(node: Node) -> bool:
        a, b = node.inputs
        for acc, bias in ((a, b), (b, a)):
            if isinstance(bias, Constant) and isinstance(bias.data, QTensor) and isinstance(acc.data, QTensor) \
                    and acc.data._pending() and acc.data._lazy.get("bias_q") is None \
                    and len(bias.data.shape) == 1 and bias.data.shape[0] == acc.data._lazy["N"]:
                return True
        return False


class PendingOutputs:
    """Handle returned by `QModel.submit`: `.result()` waits for this submission's device->host copies and
    returns the host arrays (valid until the submission after next reuses the pinned buffers)."""

    def __init__(self, outs, event):
        self._outs, self._event = outs, event

    def result(self) -> list:
        if self._event is None:
            return list(self._outs)
        self._event.synchronize()
        return [o.numpy() for o in self._outs]


def first_not_in(name: str, table: dict) -> bool:
    return name not in table


def _zp_int(zero_point):
    return None if zero_point is None else int(np.asarray(zero_point).reshape(-1)[0])
