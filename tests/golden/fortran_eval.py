"""Machine evaluation of straight-line Fortran loop bodies of the REFERENCE (golden-vector generator only).

The reference's per-cell arithmetic lives in Fortran sources that no compiler in this image can build.
Its hot loops are plain assignment statements over fp64 scalars and small fixed arrays, which map one to
one onto Python floats (IEEE binary64, same left-to-right evaluation of equal-precedence operators, no
FMA contraction, no extended precision).  `translate()` turns the text of a loop body -- taken verbatim
from the reference file at generation time, never stored in this repo -- into Python source; `run()` executes it
on given inputs.  Used by tests/golden/make_golden_fortran.py to pin the CPU oracle to the reference's own
source text.

Supported subset: assignments, `do v = a, b` / `enddo` counted loops, `do while (...)`, block and one-line
`if` with `elseif` / `else`, the dotted relational and logical operators, `write` (ignored), `stop` (raises);
`&` continuations, `!` comments and `!$omp` lines are dropped; `d0`-style exponents, dsqrt/dabs/dble
intrinsics; `x**2.0d0` becomes the correctly rounded square x*x (what a Fortran compiler emits), other
powers go to libm pow like the C oracle.  Array references are rewritten by a caller-supplied table:
per-cell arrays f(alpha,i,j,k) -> f[alpha], fields rho(i,j,k) -> rho, locals m(3) -> m[3], and `full`
arrays keep every index: f_post(alpha,i-1,j) -> f_post[(alpha,i-1,j)].
"""
import math
import re


# ---- guarded evaluation ---------------------------------------------------------------------------------------------
# The text that reaches eval()/exec() below is produced from files under /root/reference -- untrusted content.  The
# translated source is therefore parsed first and rejected unless it consists of plain arithmetic, indexing, assignments,
# loops and calls to the whitelisted names in its namespace: no attribute access, no imports, no lambdas, no dunder names,
# and it runs without Python's builtins.
import ast as _ast

_ALLOWED_NODES = (
    _ast.Module, _ast.Expression, _ast.Expr, _ast.Assign, _ast.AugAssign, _ast.For, _ast.While, _ast.If, _ast.IfExp, _ast.Break,
    _ast.Continue, _ast.Pass, _ast.BinOp, _ast.UnaryOp, _ast.BoolOp, _ast.Compare, _ast.Call, _ast.Name, _ast.Constant,
    _ast.Subscript, _ast.Tuple, _ast.List, _ast.Load, _ast.Store, _ast.Slice, _ast.Add, _ast.Sub, _ast.Mult, _ast.Div, _ast.FloorDiv,
    _ast.Mod, _ast.Pow, _ast.USub, _ast.UAdd, _ast.Not, _ast.And, _ast.Or, _ast.Eq, _ast.NotEq, _ast.Lt, _ast.LtE, _ast.Gt, _ast.GtE,
    _ast.FunctionDef, _ast.arguments, _ast.arg, _ast.Return, _ast.keyword, _ast.Global, _ast.Nonlocal, _ast.Raise)


def _checked(src, mode, what):
    tree = _ast.parse(src, what, mode)
    for node in _ast.walk(tree):
        if not isinstance(node, _ALLOWED_NODES):
            raise ValueError(f"{what}: refusing to evaluate reference-derived source containing {type(node).__name__}")
        if isinstance(node, _ast.Name) and node.id.startswith("_"):
            raise ValueError(f"{what}: refusing the name {node.id!r}")
        if isinstance(node, _ast.Call) and not isinstance(node.func, _ast.Name):
            raise ValueError(f"{what}: only calls to plain names are allowed")
    return compile(tree, what, mode)


# the only builtins translated source gets: pure functions and the exception the translation of `stop` raises
_SAFE_BUILTINS = {"dict": dict, "range": range, "abs": abs, "min": min, "max": max, "float": float, "int": int, "len": len,
                  "RuntimeError": RuntimeError}


def safe_eval(expr, ns=None, what="<reference expression>"):
    ns = ns if ns is not None else {}
    ns.setdefault("__builtins__", dict(_SAFE_BUILTINS))
    return eval(_checked(expr.strip(), "eval", what), ns)


def safe_compile(src, what="<reference source>"):
    """checked once, executed many times (safe_exec accepts the result)"""
    return _checked(src, "exec", what)


def safe_exec(src, ns, what="<reference source>"):
    ns.setdefault("__builtins__", dict(_SAFE_BUILTINS))
    exec(_checked(src, "exec", what) if isinstance(src, str) else src, ns)



def _logical_lines(text):
    out, cur = [], ""
    for raw in text.splitlines():
        line = raw.split("!")[0].rstrip()          # comments (also drops !$omp directives)
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        out.append(cur + s)
        cur = ""
    if cur:
        out.append(cur)
    return out


_NUM = re.compile(r"(?<![\w.])(\d+\.?\d*|\.\d+)[dD]([+-]?\d+)")


def _numbers(s):
    return _NUM.sub(lambda m: f"{m.group(1)}e{m.group(2)}" if m.group(2) not in ("0", "+0") else
                    (m.group(1) if "." in m.group(1) else m.group(1) + ".0"), s)


def _split_args(s):
    """split 'a, b(c,d), e' at top-level commas"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _rewrite_refs(s, cell_arrays, fields, local_arrays, full_arrays=()):
    """rewrite NAME(args) for known names; innermost-first so nested references work"""
    names = {**{n: "cell" for n in cell_arrays}, **{n: "field" for n in fields}, **{n: "local" for n in local_arrays},
             **{n: "full" for n in full_arrays}}
    if not names:
        return s
    pat = re.compile(r"\b(" + "|".join(sorted(map(re.escape, names), key=len, reverse=True)) + r")\s*\(([^()]*)\)")
    while True:
        m = pat.search(s)
        if not m:
            return s
        name, args = m.group(1), _split_args(m.group(2))
        kind = names[name]
        if kind == "cell":
            rep = f"{name}__[{args[0]}]"            # f(alpha,i,j,k) -> f__[alpha]
        elif kind == "field":
            rep = f"{name}__"                       # rho(i,j,k) -> rho__
        elif kind == "full":
            rep = f"{name}__[{'#'.join(args)}]" if len(args) > 1 else f"{name}__[{args[0]}]"   # '#' = protected comma
        else:
            rep = f"{name}__[{args[0]}]"
        s = s[:m.start()] + rep + s[m.end():]


_OPS = [(".lt.", " < "), (".le.", " <= "), (".gt.", " > "), (".ge.", " >= "), (".eq.", " == "), (".ne.", " != "),
        (".and.", " and "), (".or.", " or "), (".not.", " not ")]


def _matching_open(s, close):
    depth = 0
    for q in range(close, -1, -1):
        depth += (s[q] == ")") - (s[q] == "(")
        if depth == 0:
            return q
    raise ValueError("unbalanced parentheses: " + s)


def _powers(s):
    """x**2.0 -> sq(x) (the correctly rounded square a compiler emits); other powers -> pow(x, e)"""
    while "**" in s:
        k = s.index("**")
        left = s[:k].rstrip()
        if left.endswith(")") or left.endswith("]"):
            if left.endswith("]"):
                o = len(left) - 1
                depth = 0
                while True:
                    depth += (left[o] == "]") - (left[o] == "[")
                    if depth == 0:
                        break
                    o -= 1
            else:
                o = _matching_open(left, len(left) - 1)
            while o > 0 and (left[o - 1].isalnum() or left[o - 1] == "_"):
                o -= 1
        else:
            o = len(left)
            while o > 0 and (left[o - 1].isalnum() or left[o - 1] in "_."):
                o -= 1
        base = left[o:]
        m = re.match(r"\s*(\d+\.?\d*|\(.*?\))", s[k + 2:])
        expo = m.group(1)
        rest = s[k + 2 + m.end():]
        if float(safe_eval(expo)) == 2.0:
            s = left[:o] + f"square__({base})" + rest
        else:
            s = left[:o] + f"pow({base}, {float(safe_eval(expo))})" + rest
    return s


def _expr(line, tables):
    line = _rewrite_refs(line, *tables)
    for a, b in _OPS:
        line = line.replace(a, b)
    line = re.sub(r"\bdsqrt\b", "sqrt", line)
    line = re.sub(r"\bdabs\b", "abs", line)
    line = re.sub(r"\bdble\b", "float", line)
    line = re.sub(r"\breal\b", "float", line)
    return _powers(line).replace("#", ",")


def translate(text, cell_arrays=(), fields=(), local_arrays=(), full_arrays=()):
    """Fortran loop-body text -> Python source.  All names are lower-cased; rewritten names get a '__'
    suffix so they cannot collide with Python keywords or the intrinsics."""
    tables = ([a.lower() for a in cell_arrays], [a.lower() for a in fields], [a.lower() for a in local_arrays],
              [a.lower() for a in full_arrays])
    py, indent = [], 0
    emit = lambda txt: py.append("    " * indent + txt)
    for line in _logical_lines(text.lower()):
        for a, b in _OPS:                           # before the number pass: `.gt.1.0d0` hides the literal
            line = line.replace(a, b)
        line = _numbers(line)
        m = re.match(r"do\s+while\s*\((.*)\)\s*$", line)
        if m:
            emit(f"while {_expr(m.group(1), tables)}:")
            indent += 1
            continue
        m = re.match(r"do\s+(\w+)\s*=\s*(.+?)\s*,\s*(.+)$", line)
        if m:
            emit(f"for {m.group(1)} in range(int({_expr(m.group(2), tables)}), int({_expr(m.group(3), tables)}) + 1):")
            indent += 1
            continue
        if re.match(r"end\s*(do|if)$", line):
            indent -= 1
            continue
        m = re.match(r"(else\s*if|elseif|if)\s*\((.*)\)\s*then$", line)
        if m:
            if m.group(1) != "if":
                indent -= 1
            emit(("if " if m.group(1) == "if" else "elif ") + _expr(m.group(2), tables) + ":")
            indent += 1
            continue
        if line == "else":
            indent -= 1
            emit("else:")
            indent += 1
            continue
        if line.startswith("write") or line.startswith("call output") or line.startswith("call mpi_abort"):
            emit("pass" if line.startswith("write") else "raise RuntimeError('reference abort path')")
            continue
        if line in ("stop", "return"):
            emit("raise RuntimeError('reference stop')" if line == "stop" else "pass")
            continue
        m = re.match(r"if\s*\(", line)
        if m:                                       # one-line if
            c = 2 + line[2:].index("(")
            depth, q = 0, c
            while True:
                depth += (line[q] == "(") - (line[q] == ")")
                if depth == 0:
                    break
                q += 1
            emit(f"if {_expr(line[c + 1:q], tables)}: {_expr(line[q + 1:].strip(), tables)}")
            continue
        emit(_expr(line, tables))
    if indent != 0:
        raise ValueError("unbalanced block structure in translated body")
    return "\n".join(py)


class _Arr(dict):
    """Fortran-style small array with arbitrary lower bound."""

    def __missing__(self, k):
        raise KeyError(f"array element {k} read before assignment")


def run(py_src, cell_in=None, field_in=None, scalars=None, local_arrays=(), cell_out=(), field_out=()):
    """Execute translated source.  cell_in: {name: sequence}, field_in: {name: float},
    scalars: {name: number} (parameters such as snu, sq, nx).  Returns {name: list or float}."""
    ns = {"sqrt": math.sqrt, "abs": abs, "float": float, "int": int, "range": range, "square__": lambda x: x * x,
          "pow": math.pow, "atan": math.atan, "mod": lambda a, b: a % b}
    for name, seq in (cell_in or {}).items():
        ns[name.lower() + "__"] = _Arr({k: float(x) for k, x in enumerate(seq)})
    for name, val in (field_in or {}).items():
        ns[name.lower() + "__"] = float(val)
    for name in local_arrays:
        ns.setdefault(name.lower() + "__", _Arr())
    for name in cell_out:
        ns.setdefault(name.lower() + "__", _Arr())
    for name, val in (scalars or {}).items():
        ns[name.lower()] = val
    safe_exec(py_src, ns, "<reference loop body>")
    out = {}
    for name in cell_out:
        a = ns[name.lower() + "__"]
        out[name] = [a[k] for k in sorted(a)]
    for name in field_out:
        out[name] = ns[name.lower() + "__"]
    return out


def read_lines(path, first, last):
    """lines first..last (1-based, inclusive) of a reference source file"""
    with open(path, errors="replace") as fh:
        lines = fh.read().splitlines()
    return "\n".join(lines[first - 1:last])
