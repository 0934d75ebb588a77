#!/usr/bin/env python3
"""Fortran-90-subset -> C source-to-source translator (TEST INFRASTRUCTURE ONLY).

Purpose
-------
The reference (onera/Broadcast) is Fortran behind f2py and neither this container nor
the GPU box has a Fortran compiler, so the reference can never be *executed* as shipped.
This tool machine-translates the reference's own hot-path Fortran files, *where they lie*
under /root/reference (read-only), into C that gcc can build.  The output goes to the
git-ignored ``oracle/_ref/`` directory only; no reference source is ever copied into the
repository.  The resulting ``libbroadcast_ref.so`` is the closest thing to "the reference
run here": it is used to pin the hand-written oracle (``oracle/broadcast_oracle.py``), to
generate the golden fixtures under ``tests/golden/`` and as the timed CPU baseline.

It is NOT part of the product path: nothing under ``broadcast_b200/`` imports it.

Supported subset (census of the ~25 files on the hot path):
``subroutine``/``end subroutine``; ``implicit none``; ``real(8)|REAL*8|integer|character``
declarations with ``dimension``/``intent``/``pointer``/``target``; ``do``/``enddo``;
block and one-line ``if``; ``call``; scalar, array-element, array-section and whole-array
assignments; intrinsics SQRT ABS TANH MAX MIN SIGN INT FLOAT DBLE MOD EXP LOG; ``**``;
continuation ``&``; ``;`` statement separators; pointer assignment ``=>`` (ignored: the
only users, jn_match*.F90, never dereference them).

All REAL literals are double precision (every reference module is built with ``-r8`` /
``-fdefault-real-8``: srcfv/compile_rhs.py:29,33, srcfv/compile_tangent.py:28-30).
"""
from __future__ import annotations

import json
import re
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# --------------------------------------------------------------------------------------
# source normalisation
# --------------------------------------------------------------------------------------


def strip_comment(line: str) -> str:
    out = []
    q = None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def lower_outside_strings(s: str) -> str:
    out = []
    q = None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


def logical_lines(text: str) -> List[Tuple[int, str]]:
    """Join continuations, strip comments, split on ';'. Returns (first line no, stmt)."""
    res: List[Tuple[int, str]] = []
    cur = ""
    cur_no = 0
    for no, raw in enumerate(text.splitlines(), 1):
        if raw.lstrip().startswith("#"):
            continue
        line = strip_comment(raw)
        if not line.strip():
            continue
        s = line.strip()
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
            cur += " " + s
        else:
            cur = s
            cur_no = no
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        # split on ';' outside strings
        parts = []
        buf = []
        q = None
        for ch in cur:
            if q:
                buf.append(ch)
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                buf.append(ch)
            elif ch == ";":
                parts.append("".join(buf))
                buf = []
            else:
                buf.append(ch)
        parts.append("".join(buf))
        for p in parts:
            p = lower_outside_strings(p.strip())
            if p:
                res.append((cur_no, p))
        cur = ""
    return res


# --------------------------------------------------------------------------------------
# expression tokenizer / parser
# --------------------------------------------------------------------------------------

TOKEN_RE = re.compile(
    r"""\s*(?:
      (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
    | (?P<dotop>\.(?:eq|ne|lt|le|gt|ge|and|or|not|true|false)\.)
    | (?P<id>[a-z_]\w*)
    | (?P<str>'[^']*'|"[^"]*")
    | (?P<op>\*\*|=>|==|/=|<=|>=|[-+*/(),:<>=])
    )""",
    re.X,
)


def tokenize(s: str) -> List[Tuple[str, str]]:
    toks = []
    pos = 0
    s = s.strip()
    while pos < len(s):
        # a number followed by a dot-operator, e.g. "1.eq.x": regex alternation order
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {s[pos:pos+30]!r} in {s!r}")
        kind = m.lastgroup
        val = m.group(kind)
        if kind == "num":
            # guard "2.eq." style: if number ends with '.' and is followed by letters+'.'
            mm = re.match(r"(\d+)\.(eq|ne|lt|le|gt|ge|and|or)\.", s[m.start(kind):])
            if mm:
                val = mm.group(1)
                pos = m.start(kind) + len(val)
                toks.append(("num", val))
                continue
        toks.append((kind, val))
        pos = m.end()
    return toks


@dataclass
class Node:
    kind: str  # num, id, str, call (array ref or function), bin, un, colon, logical
    val: str = ""
    args: List["Node"] = field(default_factory=list)


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expect(self, val):
        k, v = self.next()
        if v != val:
            raise SyntaxError(f"expected {val!r} got {v!r} in {self.t}")

    # precedence: .or. < .and. < .not. < relational < +,- < *,/ < unary < **
    def parse(self):
        n = self.p_or()
        return n

    def p_or(self):
        n = self.p_and()
        while self.peek()[1] == ".or.":
            self.next()
            n = Node("bin", "||", [n, self.p_and()])
        return n

    def p_and(self):
        n = self.p_not()
        while self.peek()[1] == ".and.":
            self.next()
            n = Node("bin", "&&", [n, self.p_not()])
        return n

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.next()
            return Node("un", "!", [self.p_not()])
        return self.p_rel()

    REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
           "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}

    def p_rel(self):
        n = self.p_add()
        v = self.peek()[1]
        if v in self.REL:
            self.next()
            n = Node("bin", self.REL[v], [n, self.p_add()])
        return n

    def p_add(self):
        v = self.peek()[1]
        if v in ("+", "-"):
            self.next()
            n = Node("un", v, [self.p_mul()])
        else:
            n = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            n = Node("bin", op, [n, self.p_mul()])
        return n

    def p_mul(self):
        n = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            n = Node("bin", op, [n, self.p_pow()])
        return n

    def p_pow(self):
        n = self.p_primary()
        if self.peek()[1] == "**":
            self.next()
            # right associative; exponent may carry a unary sign
            v = self.peek()[1]
            if v in ("+", "-"):
                self.next()
                e = Node("un", v, [self.p_pow()])
            else:
                e = self.p_pow()
            n = Node("bin", "**", [n, e])
        return n

    def p_primary(self):
        k, v = self.next()
        if k == "num":
            return Node("num", v)
        if k == "str":
            return Node("str", v[1:-1])
        if k == "dotop" and v in (".true.", ".false."):
            return Node("num", "1" if v == ".true." else "0")
        if k == "id":
            if self.peek()[1] == "(":
                self.next()
                args = []
                if self.peek()[1] != ")":
                    while True:
                        args.append(self.p_arg())
                        if self.peek()[1] == ",":
                            self.next()
                            continue
                        break
                self.expect(")")
                return Node("call", v, args)
            return Node("id", v)
        if v == "(":
            n = self.p_or()
            self.expect(")")
            return Node("paren", "", [n])
        if v in ("+", "-"):
            return Node("un", v, [self.p_primary()])
        raise SyntaxError(f"unexpected token {v!r} in {self.t}")

    def p_arg(self):
        # array section "lo:hi", ":" , "lo:", ":hi" or plain expression
        if self.peek()[1] == ":":
            self.next()
            if self.peek()[1] in (",", ")"):
                return Node("colon", "", [None, None])
            hi = self.p_or()
            return Node("colon", "", [None, hi])
        lo = self.p_or()
        if self.peek()[1] == ":":
            self.next()
            if self.peek()[1] in (",", ")"):
                return Node("colon", "", [lo, None])
            hi = self.p_or()
            return Node("colon", "", [lo, hi])
        return lo


def parse_expr(s: str) -> Node:
    p = Parser(tokenize(s))
    n = p.parse()
    if p.i != len(p.t):
        raise SyntaxError(f"trailing tokens in {s!r}: {p.t[p.i:]}")
    return n


# --------------------------------------------------------------------------------------
# program-unit model
# --------------------------------------------------------------------------------------


@dataclass
class Var:
    name: str
    typ: str  # 'real' | 'int' | 'char'
    dims: Optional[List[Tuple[str, str]]] = None  # list of (lo, hi) expression strings
    intent: str = ""
    is_arg: bool = False
    assigned: bool = False
    is_pointer: bool = False


@dataclass
class Unit:
    name: str
    args: List[str]
    vars: Dict[str, Var]
    body: List[Tuple[int, str]]
    src: str = ""
    first_line: int = 0
    last_line: int = 0


def split_top(s: str, sep: str = ",") -> List[str]:
    out = []
    depth = 0
    buf = []
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(buf).strip())
            buf = []
        else:
            buf.append(ch)
    if buf or out:
        out.append("".join(buf).strip())
    return [o for o in out if o != ""]


DECL_RE = re.compile(r"^(real\s*\*\s*8|real\s*\(\s*8\s*\)|real\s*\(\s*kind\s*=\s*8\s*\)|double\s+precision|real|integer|character\s*\(\s*len\s*=\s*\d+\s*\)|character\*\d+|character|logical)\b(.*)$")


def parse_dims(s: str) -> List[Tuple[str, str]]:
    dims = []
    for d in split_top(s):
        parts = split_top(d, ":")
        if len(parts) == 1:
            dims.append(("1", parts[0]))
        else:
            dims.append((parts[0], parts[1]))
    return dims


def parse_decl(stmt: str, vars_: Dict[str, Var]) -> bool:
    m = DECL_RE.match(stmt)
    if not m:
        return False
    tword = m.group(1)
    rest = m.group(2).strip()
    if tword.startswith("real") or tword.startswith("double"):
        typ = "real"
    elif tword.startswith("integer") or tword.startswith("logical"):
        typ = "int"
    else:
        typ = "char"
    attrs = ""
    if "::" in rest:
        attrs, ents = rest.split("::", 1)
    else:
        # "integer i,j" style without '::' ; make sure it is not e.g. "real(8) function"
        ents = rest
    dims = None
    intent = ""
    is_pointer = False
    for a in split_top(attrs.strip().lstrip(",")):
        a = a.strip()
        if a == "pointer":
            is_pointer = True
        if a.startswith("dimension"):
            dims = parse_dims(a[a.index("(") + 1 : a.rindex(")")])
        elif a.startswith("intent"):
            intent = a[a.index("(") + 1 : a.rindex(")")].strip()
    for e in split_top(ents):
        e = e.strip()
        edims = dims
        if "(" in e:
            nm = e[: e.index("(")].strip()
            edims = parse_dims(e[e.index("(") + 1 : e.rindex(")")])
        else:
            nm = e
        if "=" in nm:
            raise SyntaxError("initialisers not supported: " + stmt)
        v = vars_.get(nm)
        if v is None:
            vars_[nm] = Var(nm, typ, edims, intent)
            v = vars_[nm]
        else:
            v.typ = typ
            v.dims = edims
            v.intent = intent
        v.is_pointer = is_pointer
    return True


def parse_units(text: str, names: Optional[List[str]] = None) -> List[Unit]:
    lines = logical_lines(text)
    units: List[Unit] = []
    i = 0
    while i < len(lines):
        no, s = lines[i]
        m = re.match(r"^subroutine\s+(\w+)\s*\((.*)\)\s*$", s)
        if not m:
            i += 1
            continue
        name = m.group(1)
        if names is not None and name not in names:
            i += 1
            continue
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        vars_: Dict[str, Var] = {}
        body = []
        i += 1
        while i < len(lines):
            no2, s2 = lines[i]
            if re.match(r"^end\s*subroutine\b", s2) or s2 == "end":
                break
            if s2.startswith("implicit ") or s2.startswith("intrinsic ") or s2.startswith("external "):
                i += 1
                continue
            if not body and parse_decl(s2, vars_):
                i += 1
                continue
            if body and DECL_RE.match(s2) and "::" in s2:
                # late declaration (Tapenade puts some after INTRINSIC lines)
                parse_decl(s2, vars_)
                i += 1
                continue
            body.append((no2, s2))
            i += 1
        for a in args:
            if a not in vars_:
                raise SyntaxError(f"{name}: argument {a} undeclared")
            vars_[a].is_arg = True
        units.append(Unit(name, args, vars_, body, first_line=no, last_line=lines[i][0] if i < len(lines) else no))
        i += 1
    return units


# --------------------------------------------------------------------------------------
# C emission
# --------------------------------------------------------------------------------------

REAL_FUNCS = {"sqrt": "sqrt", "tanh": "tanh", "exp": "exp", "log": "log", "dsqrt": "sqrt",
              "sin": "sin", "cos": "cos", "atan": "atan", "dabs": "fabs", "tan": "tan",
              "sinh": "sinh", "cosh": "cosh", "acos": "acos", "asin": "asin", "log10": "log10"}

PRELUDE = r"""/* GENERATED by oracle/f90_to_c.py from the read-only reference tree -- do not commit. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static inline double f_powi(double x, int n) {
  if (n == 2) return x * x;
  int neg = n < 0; unsigned u = neg ? (unsigned)(-n) : (unsigned)n;
  double r = 1.0, b = x;
  while (u) { if (u & 1u) r *= b; u >>= 1; if (u) b *= b; }
  return neg ? 1.0 / r : r;
}
static inline int f_ipow(int x, int n) { int r = 1; for (int q = 0; q < n; ++q) r *= x; return r; }
static inline double f_maxd(double a, double b) { return a > b ? a : b; }
static inline double f_mind(double a, double b) { return a < b ? a : b; }
static inline int f_maxi(int a, int b) { return a > b ? a : b; }
static inline int f_mini(int a, int b) { return a < b ? a : b; }
static inline double f_signd(double a, double b) { return copysign(fabs(a), b); }
static inline int f_signi(int a, int b) { int m = a < 0 ? -a : a; return b >= 0 ? m : -m; }
"""


class Emitter:
    def __init__(self, unit: Unit, all_units: Dict[str, Unit]):
        self.u = unit
        self.all = all_units
        self.out: List[str] = []
        self.ind = 1
        self.tmp = 0
        # scalar dummies that are assigned become pointers
        self.byref: set = set()

    # -- helpers ---------------------------------------------------------------------
    def emit(self, s: str):
        self.out.append("  " * self.ind + s)

    def var(self, name: str) -> Optional[Var]:
        return self.u.vars.get(name)

    def is_array(self, name: str) -> bool:
        v = self.var(name)
        return v is not None and v.dims is not None

    def cname(self, name: str) -> str:
        return name + "_"

    # -- typing ----------------------------------------------------------------------
    def typeof(self, n: Node) -> str:
        k = n.kind
        if k == "num":
            return "int" if re.fullmatch(r"\d+", n.val) else "real"
        if k == "str":
            return "char"
        if k == "id":
            v = self.var(n.val)
            if v is None:
                raise SyntaxError(f"{self.u.name}: unknown identifier {n.val}")
            return v.typ
        if k == "paren":
            return self.typeof(n.args[0])
        if k == "un":
            return "int" if n.val == "!" else self.typeof(n.args[0])
        if k == "bin":
            if n.val in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return "int"
            a, b = self.typeof(n.args[0]), self.typeof(n.args[1])
            if n.val == "**":
                return a
            return "real" if "real" in (a, b) else "int"
        if k == "call":
            if self.is_array(n.val):
                return self.var(n.val).typ
            f = n.val
            if f in REAL_FUNCS or f in ("float", "dble", "real"):
                return "real"
            if f in ("int", "nint", "floor"):
                return "int"
            if f in ("abs", "max", "min", "sign", "mod"):
                ts = [self.typeof(a) for a in n.args]
                return "real" if "real" in ts else "int"
            raise SyntaxError(f"{self.u.name}: unknown function/array {f}")
        raise SyntaxError("typeof: " + k)

    # -- expressions -----------------------------------------------------------------
    def num(self, s: str) -> str:
        s = re.sub(r"_\w+$", "", s)
        if re.fullmatch(r"\d+", s):
            return s
        s = s.replace("d", "e")
        if "e" in s:
            mant, ex = s.split("e")
            if "." not in mant:
                mant += ".0"
            elif mant.endswith("."):
                mant += "0"
            return mant + "e" + ex
        if s.endswith("."):
            s += "0"
        if s.startswith("."):
            s = "0" + s
        return s

    def ex(self, n: Node, secmap: Optional[dict] = None) -> str:
        k = n.kind
        if k == "num":
            return self.num(n.val)
        if k == "id":
            v = self.var(n.val)
            if v is None:
                raise SyntaxError(f"{self.u.name}: unknown identifier {n.val}")
            if v.dims is not None:
                if secmap is not None and "flat" in secmap:
                    return f"{n.val}_p[{secmap['flat']}]"
                raise SyntaxError(f"{self.u.name}: bare array {n.val} in scalar expression")
            return self.cname(n.val)
        if k == "paren":
            return "(" + self.ex(n.args[0], secmap) + ")"
        if k == "un":
            return "(" + n.val + self.ex(n.args[0], secmap) + ")"
        if k == "bin":
            a, b = n.args
            if n.val == "**":
                ta, tb = self.typeof(a), self.typeof(b)
                if tb == "int":
                    if ta == "int":
                        return f"f_ipow({self.ex(a, secmap)}, {self.ex(b, secmap)})"
                    return f"f_powi({self.ex(a, secmap)}, {self.ex(b, secmap)})"
                return f"pow({self.ex(a, secmap)}, {self.ex(b, secmap)})"
            if n.val in ("==", "!=") and (self.typeof(a) == "char" or self.typeof(b) == "char"):
                lhs, rhs = (a, b) if b.kind == "str" else (b, a)
                cmp = f"(strncmp({self.ex(lhs, secmap)}, \"{rhs.val}\", {len(rhs.val)}) == 0)"
                return cmp if n.val == "==" else "(!" + cmp + ")"
            return "(" + self.ex(a, secmap) + " " + n.val + " " + self.ex(b, secmap) + ")"
        if k == "call":
            if self.is_array(n.val):
                return self.aref(n, secmap)
            f = n.val
            args = [self.ex(a, secmap) for a in n.args]
            t = self.typeof(n)
            if f in REAL_FUNCS:
                return f"{REAL_FUNCS[f]}({args[0]})"
            if f in ("float", "dble", "real"):
                return f"((double)({args[0]}))"
            if f == "int":
                return f"((int)({args[0]}))"
            if f == "abs":
                return f"fabs({args[0]})" if t == "real" else f"abs({args[0]})"
            if f in ("max", "min"):
                fn = ("f_max" if f == "max" else "f_min") + ("d" if t == "real" else "i")
                r = args[0]
                for a in args[1:]:
                    r = f"{fn}({r}, {a})"
                return r
            if f == "sign":
                return (f"f_signd({args[0]}, {args[1]})" if t == "real" else f"f_signi({args[0]}, {args[1]})")
            if f == "mod":
                return f"fmod({args[0]}, {args[1]})" if t == "real" else f"(({args[0]}) % ({args[1]}))"
            raise SyntaxError(f"{self.u.name}: unknown function {f}")
        if k == "str":
            return '"' + n.val + '"'
        raise SyntaxError("ex: " + k)

    def aref(self, n: Node, secmap: Optional[dict]) -> str:
        v = self.var(n.val)
        if len(n.args) != len(v.dims):
            raise SyntaxError(f"{self.u.name}: rank mismatch on {n.val}")
        idx = []
        sec_no = 0
        for d, a in enumerate(n.args):
            if a.kind == "colon":
                if secmap is None or "loops" not in secmap:
                    raise SyntaxError(f"{self.u.name}: array section of {n.val} outside section assignment")
                lo = self.ex(a.args[0]) if a.args[0] is not None else f"{n.val}_l{d+1}"
                idx.append(f"(({lo}) + {secmap['loops'][sec_no]})")
                sec_no += 1
            else:
                idx.append(self.ex(a, secmap))
        return f"{n.val}_({', '.join(idx)})"

    # -- statements ------------------------------------------------------------------
    def assign(self, lhs: str, rhs: str, no: int):
        L = parse_expr(lhs)
        R = parse_expr(rhs)
        if L.kind == "id":
            v = self.var(L.val)
            if v is None:
                raise SyntaxError(f"{self.u.name}:{no}: unknown {L.val}")
            v.assigned = True
            if v.dims is not None:
                # whole array assignment
                self.emit(f"for (long q_ = 0; q_ < {L.val}_tot; ++q_) {L.val}_p[q_] = {self.ex(R, {'flat': 'q_'})};")
                return
            self.emit(f"{self.cname(L.val)} = {self.ex(R)};")
            return
        if L.kind == "call" and self.is_array(L.val):
            self.var(L.val).assigned = True
            v = self.var(L.val)
            secs = [(d, a) for d, a in enumerate(L.args) if a.kind == "colon"]
            if not secs:
                self.emit(f"{self.aref(L, None)} = {self.ex(R)};")
                return
            loops = []
            self.emit("{")
            self.ind += 1
            for sn, (d, a) in enumerate(secs):
                lo = self.ex(a.args[0]) if a.args[0] is not None else f"{L.val}_l{d+1}"
                hi = self.ex(a.args[1]) if a.args[1] is not None else f"({L.val}_l{d+1} + {L.val}_n{d+1} - 1)"
                lv = f"s{sn}_"
                loops.append(lv)
                self.emit(f"const long {lv}n = (long)({hi}) - (long)({lo}) + 1;")
            # innermost = first section (column-major friendly)
            for sn in reversed(range(len(secs))):
                lv = loops[sn]
                self.emit(f"for (long {lv} = 0; {lv} < {lv}n; ++{lv})")
                self.ind += 1
            sm = {"loops": loops}
            self.emit(f"{self.aref(L, sm)} = {self.ex(R, sm)};")
            self.ind -= len(secs)
            self.ind -= 1
            self.emit("}")
            return
        raise SyntaxError(f"{self.u.name}:{no}: bad assignment target {lhs}")

    def call(self, stmt: str, no: int):
        m = re.match(r"^call\s+(\w+)\s*\((.*)\)\s*$", stmt)
        if not m:
            raise SyntaxError(f"{self.u.name}:{no}: bad call {stmt}")
        callee = m.group(1)
        if callee not in self.all:
            raise SyntaxError(f"{self.u.name}:{no}: call to untranslated {callee}")
        cu = self.all[callee]
        args = split_top(m.group(2))
        cargs = []
        for a, formal in zip(args, cu.args):
            fv = cu.vars[formal]
            n = parse_expr(a)
            if fv.dims is not None:
                if n.kind != "id" or not self.is_array(n.val):
                    raise SyntaxError(f"{self.u.name}:{no}: array actual expected for {formal}")
                self.var(n.val).assigned = True
                cargs.append(f"{n.val}_p")
            elif fv.typ == "char":
                cargs.append(self.ex(n))
            else:
                cargs.append(self.ex(n))
        self.emit(f"{callee}({', '.join(cargs)});")

    def find_matching_paren(self, s: str, start: int) -> int:
        depth = 0
        for i in range(start, len(s)):
            if s[i] == "(":
                depth += 1
            elif s[i] == ")":
                depth -= 1
                if depth == 0:
                    return i
        raise SyntaxError("unbalanced parens: " + s)

    def stmt(self, no: int, s: str):
        if s.startswith("call "):
            self.call(s, no)
            return
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
        if m:
            var = m.group(1)
            parts = split_top(m.group(2))
            lo = self.ex(parse_expr(parts[0]))
            hi = self.ex(parse_expr(parts[1]))
            self.var(var).assigned = True
            cv = self.cname(var)
            self.tmp += 1
            t = self.tmp
            if len(parts) == 3:
                st = self.ex(parse_expr(parts[2]))
                self.emit(f"{{ const int hi{t}_ = {hi}; const int st{t}_ = {st};")
                self.emit(f"for ({cv} = {lo}; (st{t}_ > 0) ? ({cv} <= hi{t}_) : ({cv} >= hi{t}_); {cv} += st{t}_) {{")
            else:
                self.emit(f"{{ const int hi{t}_ = {hi};")
                self.emit(f"for ({cv} = {lo}; {cv} <= hi{t}_; ++{cv}) {{")
            self.ind += 1
            return
        if re.match(r"^end\s*do$", s):
            self.ind -= 1
            self.emit("} }")
            return
        if re.match(r"^else\s*if\b", s):
            p0 = s.index("(")
            p1 = self.find_matching_paren(s, p0)
            cond = self.ex(parse_expr(s[p0 + 1 : p1]))
            self.ind -= 1
            self.emit(f"}} else if ({cond}) {{")
            self.ind += 1
            return
        if s == "else":
            self.ind -= 1
            self.emit("} else {")
            self.ind += 1
            return
        if re.match(r"^end\s*if$", s):
            self.ind -= 1
            self.emit("}")
            return
        if re.match(r"^if\s*\(", s):
            p0 = s.index("(")
            p1 = self.find_matching_paren(s, p0)
            cond = self.ex(parse_expr(s[p0 + 1 : p1]))
            rest = s[p1 + 1 :].strip()
            if rest == "then":
                self.emit(f"if ({cond}) {{")
                self.ind += 1
            else:
                self.emit(f"if ({cond}) {{")
                self.ind += 1
                self.stmt(no, rest)
                self.ind -= 1
                self.emit("}")
            return
        if s in ("return",):
            self.emit("goto done_;")
            return
        if s == "continue":
            return
        if "=>" in s:
            # scalar pointer association  p => t  (jn_match_geom.F90:46-52 dereferences them)
            lhs, rhs = [t.strip() for t in s.split("=>", 1)]
            self.emit(f"{lhs}_q = &{self.cname(rhs)};")
            return
        # assignment: find top-level '=' that is not part of ==, /=, <=, >=
        depth = 0
        for i, ch in enumerate(s):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if s[i + 1 : i + 2] == "=" or s[i - 1] in "/<>=":
                    continue
                self.assign(s[:i].strip(), s[i + 1 :].strip(), no)
                return
        raise SyntaxError(f"{self.u.name}:{no}: unsupported statement: {s}")

    # -- unit ------------------------------------------------------------------------
    def signature(self) -> str:
        ps = []
        for a in self.u.args:
            v = self.u.vars[a]
            if v.dims is not None:
                ps.append(("double" if v.typ == "real" else "int") + f" *{a}_p")
            elif v.typ == "char":
                ps.append(f"const char *{a}_")
            elif a in self.byref:
                ps.append(("double" if v.typ == "real" else "int") + f" *{a}_r")
            else:
                ps.append(("double" if v.typ == "real" else "int") + f" {a}_")
        return f"void {self.u.name}({', '.join(ps)})"

    def run(self) -> str:
        u = self.u
        # pass 1: emit body into a scratch buffer to discover assigned scalar dummies
        body_out: List[str] = []
        saved = self.out
        self.out = body_out
        for no, s in u.body:
            try:
                self.stmt(no, s)
            except SyntaxError as e:
                raise SyntaxError(f"{u.name}: line {no}: {s}\n   {e}")
        self.out = saved
        for a in u.args:
            v = u.vars[a]
            if v.dims is None and v.typ != "char" and v.assigned:
                self.byref.add(a)
        hdr = [self.signature() + " {"]
        undef = []
        # scalars
        for name, v in u.vars.items():
            if v.dims is None:
                if v.is_arg:
                    if name in self.byref:
                        hdr.append(f"#define {name}_ (*{name}_r)")
                        undef.append(f"{name}_")
                    continue
                if v.is_pointer:
                    ct = 'double' if v.typ == 'real' else 'int'
                    hdr.append(f"  {ct} {name}_dummy = 0; {ct} *{name}_q = &{name}_dummy;")
                    hdr.append(f"#define {name}_ (*{name}_q)")
                    undef.append(f"{name}_")
                elif v.typ == "char":
                    hdr.append(f"  const char *{name}_ = \"\"; (void){name}_;")
                else:
                    hdr.append(f"  {'double' if v.typ == 'real' else 'int'} {name}_ = 0; (void){name}_;")
        # arrays (bounds depend on scalar dummies only)
        for name, v in u.vars.items():
            if v.dims is None:
                continue
            ctype = "double" if v.typ == "real" else "int"
            tot = []
            for d, (lo, hi) in enumerate(v.dims, 1):
                hdr.append(f"  const long {name}_l{d} = {self.ex(parse_expr(lo))};")
                hdr.append(f"  const long {name}_n{d} = (long)({self.ex(parse_expr(hi))}) - {name}_l{d} + 1; (void){name}_n{d};")
                tot.append(f"{name}_n{d}")
            hdr.append(f"  const long {name}_tot = {' * '.join(tot)}; (void){name}_tot;")
            if not v.is_arg:
                hdr.append(f"  {ctype} *{name}_p = ({ctype} *)calloc((size_t)({name}_tot > 0 ? {name}_tot : 1), sizeof({ctype}));")
            idxs = [f"i{d}" for d in range(1, len(v.dims) + 1)]
            # column-major offset, built inside-out
            expr = f"((long)(i{len(v.dims)}) - {name}_l{len(v.dims)})"
            for d in reversed(range(1, len(v.dims))):
                expr = f"(((long)(i{d}) - {name}_l{d}) + {name}_n{d} * {expr})"
            hdr.append(f"#define {name}_({', '.join(idxs)}) {name}_p[{expr}]")
            undef.append(f"{name}_")
        tail = ["  goto done_; done_: ;"]
        for name, v in u.vars.items():
            if v.dims is not None and not v.is_arg:
                tail.append(f"  free({name}_p);")
        for m in undef:
            tail.append(f"#undef {m}")
        tail.append("}")
        return "\n".join(hdr + body_out + tail) + "\n"


def manifest_entry(u: Unit) -> dict:
    args = []
    for a in u.args:
        v = u.vars[a]
        args.append({"name": a, "type": v.typ, "dims": v.dims, "intent": v.intent})
    return {"name": u.name, "args": args, "first_line": u.first_line, "last_line": u.last_line}


def translate(files: List[Tuple[str, Optional[List[str]]]]) -> Tuple[str, List[dict]]:
    """files: list of (path, [subroutine names] or None for all). Returns (C text, manifest)."""
    units: List[Unit] = []
    for path, names in files:
        with open(path, "r", errors="replace") as fh:
            us = parse_units(fh.read(), names)
        for u in us:
            u.src = path
            units.append(u)
    allu = {u.name: u for u in units}
    # callees first
    order: List[Unit] = []
    seen = set()

    def visit(u: Unit):
        if u.name in seen:
            return
        seen.add(u.name)
        for _, s in u.body:
            m = re.match(r"^(?:if\s*\(.*\)\s*)?call\s+(\w+)", s)
            if m and m.group(1) in allu:
                visit(allu[m.group(1)])
        order.append(u)

    for u in units:
        visit(u)
    chunks = [PRELUDE]
    manifest = []
    for u in order:
        chunks.append(f"/* ---- {u.name}: translated from {u.src}:{u.first_line}-{u.last_line} ---- */")
        chunks.append(Emitter(u, allu).run())
        me = manifest_entry(u)
        me["src"] = u.src
        manifest.append(me)
    return "\n".join(chunks), manifest


if __name__ == "__main__":
    spec = json.load(open(sys.argv[1]))
    c, man = translate([(f["path"], f.get("subs")) for f in spec["files"]])
    open(sys.argv[2], "w").write(c)
    json.dump(man, open(sys.argv[3], "w"), indent=1)
