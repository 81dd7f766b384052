"""Minimal stand-in for the `gin-config` package (pinned by the reference at
requirements.txt:4, absent from this image and from /opt/wheelhouse).

Only the surface the hot path's drop-in callers use is implemented
(SURVEY.md §5 "Config / flags", §8(c)):

* ``@gin.configurable`` on classes and functions (bare or with arguments),
* ``gin.external_configurable(obj, name=None, module=None)``,
* ``with gin.config_scope("noise_synth"):`` (scoped bindings),
* ``gin.parse_config_file(path)`` / ``gin.parse_config(text)`` with macros
  (``x = 1``), macro references (``%x``), configurable references (``@Name`` /
  ``@Name()``), scoped bindings (``scope/Class.arg = v``), ``include '...'``
  and ``import a.b`` lines,
* ``gin.constant``, ``gin.bind_parameter``, ``gin.query_parameter``,
  ``gin.clear_config``, ``gin.REQUIRED``.

Semantics follow gin's documented behaviour: a binding supplies a value for a
parameter the caller did not pass; explicitly passed arguments always win; a
scoped binding applies only while its scope is active and beats an unscoped one.
"""
from __future__ import annotations

import ast
import contextlib
import functools
import importlib
import inspect
import os
import threading
from typing import Any, Callable, Dict, List, Optional, Tuple

__all__ = [
    "configurable", "external_configurable", "config_scope", "parse_config_file",
    "parse_config", "parse_config_files_and_bindings", "constant", "bind_parameter",
    "query_parameter", "clear_config", "REQUIRED", "config_str", "operative_config_str",
    "add_config_file_search_path",
]


class _Required:
    def __repr__(self):
        return "gin.REQUIRED"


REQUIRED = _Required()

# name -> wrapped callable (both short "Name" and qualified "module.Name")
_REGISTRY: Dict[str, Callable] = {}
# (scope, selector, arg) -> value ; scope == "" for unscoped bindings
_BINDINGS: Dict[Tuple[str, str, str], Any] = {}
_MACROS: Dict[str, Any] = {}
_CONSTANTS: Dict[str, Any] = {}
_OPERATIVE: Dict[Tuple[str, str, str], Any] = {}
_SEARCH_PATHS: List[str] = [""]
_tls = threading.local()


def _scopes() -> List[str]:
    if not hasattr(_tls, "scopes"):
        _tls.scopes = []
    return _tls.scopes


class _MacroRef:
    def __init__(self, name):
        self.name = name

    def resolve(self):
        if self.name in _MACROS:
            return _resolve(_MACROS[self.name])
        if self.name in _CONSTANTS:
            return _CONSTANTS[self.name]
        raise ValueError("gin: undefined macro %%%s" % self.name)


class _ConfigurableRef:
    def __init__(self, name, call, scope=""):
        self.name, self.call, self.scope = name, call, scope

    def resolve(self):
        fn = _lookup(self.name)
        if not self.call:
            return fn
        with contextlib.ExitStack() as stack:
            for s in [p for p in self.scope.split("/") if p]:
                stack.enter_context(config_scope(s))
            return fn()


def _resolve(value):
    if isinstance(value, (_MacroRef, _ConfigurableRef)):
        return value.resolve()
    if isinstance(value, list):
        return [_resolve(v) for v in value]
    if isinstance(value, tuple):
        return tuple(_resolve(v) for v in value)
    if isinstance(value, dict):
        return {_resolve(k): _resolve(v) for k, v in value.items()}
    return value


def _lookup(name: str) -> Callable:
    if name in _REGISTRY:
        return _REGISTRY[name]
    # allow partially-qualified selectors: match on the trailing components
    hits = {id(v): v for k, v in _REGISTRY.items() if k.endswith("." + name)}
    if len(hits) == 1:
        return next(iter(hits.values()))
    raise ValueError("gin: no configurable named %r" % name)


def _selector_matches(selector: str, names: Tuple[str, ...]) -> bool:
    return any(selector == n or n.endswith("." + selector) for n in names)


def _find_binding(names: Tuple[str, ...], arg: str):
    """Most specific active binding for (configurable, arg) or (False, None)."""
    active = _scopes()
    best = None
    for (scope, selector, a), value in _BINDINGS.items():
        if a != arg or not _selector_matches(selector, names):
            continue
        if scope:
            parts = scope.split("/")
            # the binding's scope path must be a suffix-aligned subsequence of the active stack
            it = iter(active)
            if not all(p in it for p in parts):
                continue
            rank = len(parts)
        else:
            rank = 0
        if best is None or rank > best[0]:
            best = (rank, (scope, selector, a), value)
    if best is None:
        return False, None
    _OPERATIVE[best[1]] = best[2]
    return True, _resolve(best[2])


def _wrap(fn: Callable, names: Tuple[str, ...], skip_self: bool) -> Callable:
    sig = inspect.signature(fn)
    params = list(sig.parameters.values())
    if skip_self:
        params = params[1:]
    positional = [p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    injectable = [p for p in params if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        n_pos = len(args) - (1 if skip_self else 0)
        given = {p.name for p in positional[:n_pos]} | set(kwargs)
        for p in injectable:
            if p.name in given:
                if kwargs.get(p.name, None) is REQUIRED:
                    found, value = _find_binding(names, p.name)
                    if not found:
                        raise RuntimeError("gin: required binding missing for %s.%s" % (names[0], p.name))
                    kwargs[p.name] = value
                continue
            found, value = _find_binding(names, p.name)
            if found:
                kwargs[p.name] = value
        return fn(*args, **kwargs)

    wrapper.__gin_names__ = names
    return wrapper


def _register(obj, name: Optional[str], module: Optional[str], subclass: bool):
    short = name or obj.__name__
    mod = module if module is not None else getattr(obj, "__module__", None)
    names = (short,) + ((mod + "." + short,) if mod else ())
    if inspect.isclass(obj):
        if subclass:
            target = type(obj.__name__, (obj,), {"__module__": obj.__module__, "__doc__": obj.__doc__})
            base_init = obj.__init__

            def __init__(self, *a, **k):  # noqa: N807
                base_init(self, *a, **k)

            __init__.__signature__ = inspect.signature(base_init)
            target.__init__ = _wrap(__init__, names, skip_self=True)
        else:
            target = obj
            target.__init__ = _wrap(obj.__init__, names, skip_self=True)
        wrapped = target
    else:
        wrapped = _wrap(obj, names, skip_self=False)
    for n in names:
        _REGISTRY[n] = wrapped
    return wrapped


def configurable(name_or_fn=None, module=None, allowlist=None, denylist=None, whitelist=None, blacklist=None):
    """``@gin.configurable`` (reference use: models/neural_waveshaping.py:16,29; modules/*.py)."""
    if callable(name_or_fn):
        return _register(name_or_fn, None, module, subclass=False)

    def deco(obj):
        return _register(obj, name_or_fn, module, subclass=False)

    return deco


def external_configurable(fn_or_cls, name=None, module=None, allowlist=None, denylist=None,
                          whitelist=None, blacklist=None):
    """``gin.external_configurable`` (reference use: models/neural_waveshaping.py:13-14)."""
    return _register(fn_or_cls, name, module, subclass=True)


@contextlib.contextmanager
def config_scope(name_or_scope):
    """``with gin.config_scope('noise_synth')`` (reference: models/neural_waveshaping.py:57)."""
    stack = _scopes()
    if name_or_scope is None or name_or_scope == "":
        saved = list(stack)
        stack.clear()
        try:
            yield []
        finally:
            stack.extend(saved)
        return
    parts = list(name_or_scope) if isinstance(name_or_scope, (list, tuple)) else \
        [p for p in str(name_or_scope).split("/") if p]
    stack.extend(parts)
    try:
        yield list(stack)
    finally:
        del stack[len(stack) - len(parts):]


def constant(name: str, value: Any):
    _CONSTANTS[name] = value
    return value


def bind_parameter(binding_key: str, value: Any):
    scope, selector, arg = _split_key(binding_key)
    _BINDINGS[(scope, selector, arg)] = value


def query_parameter(binding_key: str):
    if binding_key.startswith("%"):
        return _MacroRef(binding_key[1:]).resolve()
    scope, selector, arg = _split_key(binding_key)
    key = (scope, selector, arg)
    if key not in _BINDINGS:
        raise ValueError("gin: no binding for %r" % binding_key)
    return _resolve(_BINDINGS[key])


def clear_config(clear_constants: bool = False):
    _BINDINGS.clear()
    _MACROS.clear()
    _OPERATIVE.clear()
    if clear_constants:
        _CONSTANTS.clear()


def add_config_file_search_path(path: str):
    _SEARCH_PATHS.append(path)


def _split_key(key: str) -> Tuple[str, str, str]:
    key = key.strip()
    scope = ""
    if "/" in key:
        scope, key = key.rsplit("/", 1)
    selector, arg = key.rsplit(".", 1)
    return scope.strip("/"), selector.strip(), arg.strip()


class _ValueParser(ast.NodeTransformer):
    pass


def _parse_value(text: str):
    """Python literal with gin's %macro and @configurable extensions."""
    text = text.strip()
    # tokenise the two gin-specific forms into placeholder calls, then literal-eval the rest
    out, i, n = [], 0, len(text)
    refs: List[Any] = []
    in_str: Optional[str] = None
    while i < n:
        ch = text[i]
        if in_str:
            out.append(ch)
            if ch == "\\" and i + 1 < n:
                out.append(text[i + 1])
                i += 1
            elif ch == in_str:
                in_str = None
            i += 1
            continue
        if ch in "'\"":
            in_str = ch
            out.append(ch)
            i += 1
            continue
        if ch in "%@":
            j = i + 1
            while j < n and (text[j].isalnum() or text[j] in "_./"):
                j += 1
            ident = text[i + 1:j].rstrip(".")
            j = i + 1 + len(ident)
            if ch == "%":
                refs.append(_MacroRef(ident))
            else:
                call = text[j:j + 2] == "()"
                if call:
                    j += 2
                scope = ""
                if "/" in ident:
                    scope, ident = ident.rsplit("/", 1)
                refs.append(_ConfigurableRef(ident, call, scope))
            out.append("__gin_ref__[%d]" % (len(refs) - 1))
            i = j
            continue
        out.append(ch)
        i += 1
    expr = "".join(out)
    tree = ast.parse(expr, mode="eval")

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name) and node.value.id == "__gin_ref__":
            return refs[ev(node.slice)]
        if isinstance(node, ast.Constant):
            return node.value
        if isinstance(node, ast.List):
            return [ev(e) for e in node.elts]
        if isinstance(node, ast.Tuple):
            return tuple(ev(e) for e in node.elts)
        if isinstance(node, ast.Dict):
            return {ev(k): ev(v) for k, v in zip(node.keys, node.values)}
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else +v
        if isinstance(node, ast.Name) and node.id in ("True", "False", "None"):
            return {"True": True, "False": False, "None": None}[node.id]
        raise ValueError("gin: unsupported value syntax: %r" % text)

    return ev(tree)


def _logical_lines(text: str):
    buf, depth = "", 0
    for raw in text.splitlines():
        line = _strip_comment(raw)
        if not line.strip() and depth == 0:
            continue
        buf = (buf + " " + line) if buf else line
        depth = _bracket_depth(buf)
        if depth <= 0 and not buf.rstrip().endswith("\\"):
            yield buf.strip()
            buf, depth = "", 0
    if buf.strip():
        yield buf.strip()


def _strip_comment(line: str) -> str:
    in_str = None
    for i, ch in enumerate(line):
        if in_str:
            if ch == in_str and line[i - 1] != "\\":
                in_str = None
        elif ch in "'\"":
            in_str = ch
        elif ch == "#":
            return line[:i]
    return line


def _bracket_depth(s: str) -> int:
    depth, in_str = 0, None
    for i, ch in enumerate(s):
        if in_str:
            if ch == in_str and s[i - 1] != "\\":
                in_str = None
        elif ch in "'\"":
            in_str = ch
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
    return depth


def parse_config(bindings, skip_unknown=False):
    if isinstance(bindings, (list, tuple)):
        bindings = "\n".join(bindings)
    _parse_text(bindings or "", base_dir="")


def _parse_text(text: str, base_dir: str):
    for line in _logical_lines(text):
        if line.startswith("include "):
            inc = ast.literal_eval(line[len("include "):].strip())
            parse_config_file(_find_file(inc, base_dir))
            continue
        if line.startswith("import "):
            try:
                importlib.import_module(line[len("import "):].strip())
            except ImportError:
                pass
            continue
        if "=" not in line:
            raise ValueError("gin: cannot parse line %r" % line)
        lhs, rhs = line.split("=", 1)
        lhs = lhs.strip()
        value = _parse_value(rhs)
        if "." not in lhs.rsplit("/", 1)[-1]:
            _MACROS[lhs] = value          # `sample_rate = 16000`  (gin/models/newt.gin:1)
        else:
            _BINDINGS[_split_key(lhs)] = value


def _find_file(path: str, base_dir: str) -> str:
    if os.path.isabs(path) and os.path.exists(path):
        return path
    for root in [base_dir] + _SEARCH_PATHS:
        cand = os.path.join(root, path) if root else path
        if os.path.exists(cand):
            return cand
    raise IOError("gin: unable to open config file %r" % path)


def parse_config_file(config_file, skip_unknown=False, print_includes_and_imports=False):
    """``gin.parse_config_file`` (reference: scripts/time_forward_pass.py:26)."""
    path = _find_file(config_file, "")
    with open(path, "r") as fh:
        text = fh.read()
    _parse_text(text, base_dir=os.path.dirname(os.path.abspath(path)))


def parse_config_files_and_bindings(config_files=None, bindings=None, finalize_config=True,
                                    skip_unknown=False, print_includes_and_imports=False):
    if isinstance(config_files, str):
        config_files = [config_files]
    for f in config_files or []:
        parse_config_file(f)
    parse_config(bindings or "")


def _fmt(value) -> str:
    if isinstance(value, _MacroRef):
        return "%" + value.name
    if isinstance(value, _ConfigurableRef):
        return "@" + (value.scope + "/" if value.scope else "") + value.name + ("()" if value.call else "")
    return repr(value)


def _dump(bindings) -> str:
    lines = ["%s = %s" % (k, _fmt(v)) for k, v in _MACROS.items()]
    for (scope, selector, arg), v in bindings.items():
        lines.append("%s%s.%s = %s" % (scope + "/" if scope else "", selector, arg, _fmt(v)))
    return "\n".join(lines) + "\n"


def config_str() -> str:
    return _dump(_BINDINGS)


def operative_config_str() -> str:
    return _dump(_OPERATIVE)
