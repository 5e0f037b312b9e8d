"""`str_to_net`: build a torch module from a structure string such as

    "Linear(obs_length, 16) >> Tanh() >> Linear(16, act_length)"

(reference: net/parser.py:30-330 -- BASELINE config 4 names its policy in this notation).  The string is parsed with Python's
`ast`: a call `Name(args...)` instantiates `torch.nn.<Name>`, `>>` chains modules into a `MultiLayered`; arguments may be
literals, named constants passed as keyword arguments of `str_to_net`, or a constant indexed by a literal (`obs_shape[0]`).
Anything else is reported as a `NetParsingError` with the position in the string.
"""

from __future__ import annotations

import ast
from typing import Any, Optional

from torch import nn

from .multilayered import MultiLayered


class NetParsingError(Exception):
    def __init__(self, message: str, lineno: Optional[int] = None, col_offset: Optional[int] = None, original_error: Optional[Exception] = None):
        super().__init__()
        self.message, self.lineno, self.col_offset, self.original_error = message, lineno, col_offset, original_error

    def __str__(self) -> str:
        where = ""
        if self.lineno is not None:
            where += f" at line({self.lineno - 1})"  # the parsed text is wrapped in "(\\n ... \\n)": its first line is line 2
        if self.col_offset is not None:
            where += f" at column({self.col_offset + 1})"
        return f"{type(self).__name__}{where}: {self.message}"

    __repr__ = __str__


# names the reference resolves to its OWN layer classes (net/layers.py: unbatched recurrent cells, Clip, Bin, Slice, ...); torch.nn has
# different classes under some of these names, so they are refused rather than silently given another meaning
_REFERENCE_ONLY_LAYERS = ("RNN", "LSTM", "FeedForwardNet", "StructuredControlNet", "LocomotorNet", "Clip", "Bin", "Slice", "Round", "Apply")


def _module_class(name: str):
    if name in _REFERENCE_ONLY_LAYERS:
        raise NetParsingError(f"The layer {name!r} is one of the reference's own layer classes, which this package does not provide"
                              f" (feed-forward torch.nn modules only)")
    cls = getattr(nn, name, None)
    if not (isinstance(cls, type) and issubclass(cls, nn.Module)):
        raise NetParsingError(f"Unrecognized module class: {name!r}")
    return cls


def _value(node: ast.expr, constants: dict) -> Any:
    def fail(message: str, at: Optional[ast.AST] = None, cause: Optional[Exception] = None):
        at = at if (at is not None and hasattr(at, "lineno")) else node
        raise NetParsingError(message, at.lineno, at.col_offset, original_error=cause)

    def literal(sub: ast.expr) -> Any:
        try:
            return ast.literal_eval(sub)
        except Exception as exc:
            fail(f"When trying to parse expression, encountered: {exc!r}", sub, exc)

    def constant(name: str) -> Any:
        if name not in constants:
            fail(f"Unknown constant: {name}. Available constants: {list(constants.keys())!r}")
        return constants[name]

    if isinstance(node, ast.Name):
        return constant(node.id)
    if isinstance(node, ast.Subscript):
        if not isinstance(node.value, ast.Name):
            fail("Expression which was expected to express a simple indexing over a constant is either too complex, or is unrecognized.", node.value)
        index = node.slice
        if not isinstance(index, ast.Constant):
            fail(f"Expected a simple indexing operation, but got a {type(index).__name__}.", index)
        source = constant(node.value.id)
        try:
            return source[literal(index)]
        except NetParsingError:
            raise
        except Exception as exc:
            fail(f"When applying the indexing operation on the constant {node.value.id}, encountered: {exc!r}", cause=exc)
    return literal(node)


def _chain(left: nn.Module, right: nn.Module) -> MultiLayered:
    layers = []
    for part in (left, right):
        layers.extend(part if isinstance(part, (nn.Sequential, MultiLayered)) else [part])
    return MultiLayered(*layers)


def _module(node: ast.expr, constants: dict) -> nn.Module:
    if isinstance(node, ast.Call):
        if not isinstance(node.func, ast.Name):
            raise NetParsingError(f"Unrecognized expression of type {type(node.func)}", node.lineno, node.col_offset)
        args = [_value(a, constants) for a in node.args]
        kwargs = {k.arg: _value(k.value, constants) for k in node.keywords}
        return _module_class(node.func.id)(*args, **kwargs)
    if isinstance(node, ast.BinOp):
        if not isinstance(node.op, ast.RShift):
            raise NetParsingError("Binary operators other than '>>' are not recognized.", node.lineno, node.col_offset)
        return _chain(_module(node.left, constants), _module(node.right, constants))
    raise NetParsingError(f"Unrecognized expression of type {type(node)}", node.lineno, node.col_offset)


def str_to_net(s: str, **constants) -> nn.Module:
    """The module described by `s`; `constants` are the names usable inside it."""
    source = "(\n" + s + "\n)"
    try:
        tree = ast.parse(source, mode="eval")
    except SyntaxError as exc:
        raise NetParsingError(f"Syntax error: {exc.msg}", exc.lineno, None if exc.offset is None else exc.offset - 1, original_error=exc)
    return _module(tree.body, constants)
