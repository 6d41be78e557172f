"""f90exec -- a small Fortran-90 subset executor (TEST INFRASTRUCTURE, used only to generate golden vectors).

The reference (MoVFEM_3DMT) is Fortran and this image has no Fortran compiler, so the reference cannot be
compiled.  This module runs it anyway: it parses the reference's own source files where they lie (read-only,
nothing is copied into the repository), translates the procedures that are actually called to Python **in
memory**, and executes them with Fortran semantics:

  * default `real` / `complex` / un-suffixed real literals / `cmplx()` without kind are single precision
    (numpy float32 / complex64 scalars), `real(kind=double)` / `1.d0` are float64 / complex128; mixed-kind
    arithmetic promotes exactly as Fortran does because numpy scalars follow the same rules;
  * integers are Python ints, `/` on two integers truncates toward zero, `x**n` with integer n is a
    multiplication chain (what gfortran emits);
  * assignment converts to the declared type of the target (complex -> real takes the real part, real -> integer
    truncates);
  * arrays are column-major numpy arrays with 1-based subscripts; sections are views, so `intent(out)` array
    arguments, sequence association (`call s(a(k))` to an explicit-shape dummy) and aliasing behave as in Fortran;
    scalar dummies are copied back to the caller's variables on return;
  * modules keep their state (`save` variables, allocatables), `use ..., only:` resolves transitively.

Only what the element-assembly path of the reference needs is implemented; anything else (I/O, formats, derived
types, labels/goto) raises `Unsupported` when it is *executed*, not when it is parsed, so that large files load as
long as the called procedures stay inside the subset.  `print`/`write` statements are skipped.

Public API:   rt = Runtime([paths...]);  rt.mod('geometry').g_nx = 4;  rt.call('n_fem', 'init_n_fem', 8)
"""
import math
import re
import warnings

import numpy as np

warnings.filterwarnings("ignore", category=np.exceptions.ComplexWarning)
np.seterr(all="ignore")


class Unsupported(Exception):
    pass


class FortranStop(Exception):
    pass


# ------------------------------------------------------------------------------------------------------------
# source -> logical statements
# ------------------------------------------------------------------------------------------------------------

def _strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.lower())
    return "".join(out)


def _split_semicolons(s):
    parts, q, cur = [], None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def logical_statements(text):
    """-> list of (lineno, statement) with comments removed, continuations joined, `;` split, lower-cased."""
    out, cur, cur_no = [], "", 0
    for no, raw in enumerate(text.splitlines(), 1):
        line = _strip_comment(raw.replace("\t", " ")).rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
        else:
            cur_no = no
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        cur += s
        for st in _split_semicolons(_lower_outside_strings(cur)):
            out.append((cur_no, st))
        cur = ""
    if cur.strip():
        out.append((cur_no, _lower_outside_strings(cur)))
    return out


# ------------------------------------------------------------------------------------------------------------
# expression tokenizer / parser
# ------------------------------------------------------------------------------------------------------------

_DOTOPS = "eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv|true|false"
_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<real>(?:\d+\.(?!(?:%s)\.)\d*|\.\d+)(?:[ed][+-]?\d+)?(?:_\w+)?|\d+[ed][+-]?\d+(?:_\w+)?)
  | (?P<int>\d+(?:_\w+)?)
  | (?P<dot>\.(?:%s)\.)
  | (?P<name>[a-z_]\w*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/(),=<>:%%])
""" % (_DOTOPS, _DOTOPS), re.X)


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m:
            raise Unsupported("cannot tokenize %r at %r" % (s, s[pos:pos + 10]))
        pos = m.end()
        k = m.lastgroup
        if k != "ws":
            toks.append((k, m.group()))
    return toks


_REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
        "==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">="}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise Unsupported("expected %r at token %d of %r" % (val, self.i, self.t))

    def at_end(self):
        return self.i >= len(self.t)

    # precedence: .eqv. < .or. < .and. < .not. < relational < // < +- < */ < unary(+-) ... < **
    def expr(self):
        left = self.or_()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            left = ("bin", op, left, self.or_())
        return left

    def or_(self):
        left = self.and_()
        while self.peek()[1] == ".or.":
            self.next()
            left = ("bin", "or", left, self.and_())
        return left

    def and_(self):
        left = self.not_()
        while self.peek()[1] == ".and.":
            self.next()
            left = ("bin", "and", left, self.not_())
        return left

    def not_(self):
        if self.peek()[1] == ".not.":
            self.next()
            return ("un", "not", self.not_())
        return self.rel()

    def rel(self):
        left = self.concat()
        if self.peek()[1] in _REL:
            op = _REL[self.next()[1]]
            return ("bin", op, left, self.concat())
        return left

    def concat(self):
        left = self.add()
        while self.peek()[1] == "//":
            self.next()
            left = ("bin", "//", left, self.add())
        return left

    def add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            left = ("un", op, self.mul())
        else:
            left = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            left = ("bin", op, left, self.mul())
        return left

    def mul(self):
        left = self.pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            left = ("bin", op, left, self.pow())
        return left

    def pow(self):
        base = self.primary()
        if self.peek()[1] == "**":
            self.next()
            if self.peek()[1] in ("+", "-"):          # a**-b
                op = self.next()[1]
                return ("bin", "**", base, ("un", op, self.pow()))
            return ("bin", "**", base, self.pow())      # right associative
        return base

    def arg(self):
        """one subscript / actual argument: expr | [lo]:[hi][:step] | kw=expr"""
        k, v = self.peek()
        if k == "name" and self.peek(1)[1] == "=" :
            self.next(); self.next()
            return ("kw", v, self.expr())
        lo = None
        if self.peek()[1] != ":":
            lo = self.expr()
            if self.peek()[1] != ":":
                return lo
        self.expect(":")
        hi = step = None
        if self.peek()[1] not in (",", ")", ":"):
            hi = self.expr()
        if self.accept(":"):
            step = self.expr()
        return ("range", lo, hi, step)

    def arglist(self, close=")"):
        args = []
        if self.accept(close):
            return args
        while True:
            args.append(self.arg())
            if self.accept(close):
                return args
            self.expect(",")

    def ac_item(self):
        # implied do:  ( expr [, expr]... , var = lo, hi [, step] )
        if self.peek()[1] == "(":
            save = self.i
            try:
                self.next()
                items = [self.ac_item()]
                while self.accept(","):
                    if self.peek()[0] == "name" and self.peek(1)[1] == "=":
                        var = self.next()[1]
                        self.next()
                        lo = self.expr(); self.expect(","); hi = self.expr()
                        step = self.expr() if self.accept(",") else None
                        self.expect(")")
                        return ("ido", items, var, lo, hi, step)
                    items.append(self.ac_item())
                raise Unsupported("not an implied do")
            except Unsupported:
                self.i = save
        return self.expr()

    def primary(self):
        k, v = self.next()
        if k == "int":
            return ("int", int(v.split("_")[0]))
        if k == "real":
            return _real_literal(v)
        if k == "str":
            q = v[0]
            return ("str", v[1:-1].replace(q + q, q))
        if k == "dot":
            if v == ".true.":
                return ("log", True)
            if v == ".false.":
                return ("log", False)
            raise Unsupported("unexpected " + v)
        if k == "name":
            if self.peek()[1] == "(":
                self.next()
                args = self.arglist()
                node = ("call", v, args)
                if self.peek()[1] == "(":           # substring / chained -- not needed
                    raise Unsupported("chained reference " + v)
                return node
            if self.peek()[1] == "(/":               # name(/.../)  ==  name( (/.../) )
                self.next()
                items = self._ac_items()
                self.expect(")")
                return ("call", v, [("ac", items)])
            return ("name", v)
        if v == "(/":
            return ("ac", self._ac_items())
        if v == "(":
            e = self.expr()
            if self.accept(","):                      # complex literal (re, im)
                im = self.expr()
                self.expect(")")
                return ("cplx", e, im)
            self.expect(")")
            return ("paren", e)
        if v in ("+", "-"):
            return ("un", v, self.pow())
        raise Unsupported("unexpected token %r" % (v,))

    def _ac_items(self):
        items = []
        if self.accept("/)"):
            return items
        while True:
            items.append(self.ac_item())
            if self.accept("/)"):
                return items
            self.expect(",")


def _real_literal(v):
    kind = 4
    if "_" in v:
        v, suffix = v.split("_", 1)
        kind = 8 if suffix in ("double", "8", "dp") else 4
    if "d" in v:
        kind = 8
        v = v.replace("d", "e")
    return ("real", float(v), kind)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.at_end():
        raise Unsupported("trailing tokens in expression %r" % s)
    return e


# ------------------------------------------------------------------------------------------------------------
# declarations / program units
# ------------------------------------------------------------------------------------------------------------

class Sym:
    __slots__ = ("name", "type", "dims", "allocatable", "parameter", "init", "intent", "is_dummy")

    def __init__(self, name, type_, dims=None, allocatable=False, parameter=False, init=None, intent=None):
        self.name, self.type, self.dims = name, type_, dims
        self.allocatable, self.parameter, self.init, self.intent = allocatable, parameter, init, intent
        self.is_dummy = False

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


class Proc:
    def __init__(self, name, kind, args, result, module, lineno, rtype=None):
        self.name, self.kind, self.args, self.result, self.module = name, kind, args, result, module
        self.lineno, self.rtype = lineno, rtype
        self.uses, self.syms, self.order, self.body_lines, self.body = [], {}, [], [], None
        self.pyfunc, self.broken = None, None

    def out_scalars(self):
        """positions of scalar dummies whose value is copied back to the caller"""
        res = []
        for k, a in enumerate(self.args):
            s = self.syms.get(a)
            if s is None or (s.rank == 0 and s.intent != "in" and s.type != "proc"):
                res.append(k)
        return res


class Module:
    def __init__(self, name, path):
        self.name, self.path = name, path
        self.uses, self.syms, self.order, self.procs = [], {}, [], {}


_TYPE_RE = re.compile(r"^(integer|real|double\s*precision|complex|logical|character)\b")


def _match_paren(s, start):
    depth = 0
    q = None
    for i in range(start, len(s)):
        ch = s[i]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return i
    raise Unsupported("unbalanced parentheses in %r" % s)


def _split_top(s, sep=","):
    parts, depth, q, cur = [], 0, None, []
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "(":
            depth += 1
            cur.append(ch)
        elif ch == ")":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur).strip())
    return parts


def _parse_type_spec(s):
    """s starts with a type keyword; returns (type_code, rest)"""
    m = _TYPE_RE.match(s)
    base = m.group(1).replace(" ", "")
    rest = s[m.end():].lstrip()
    kind = None
    if rest.startswith("*"):
        m2 = re.match(r"\*\s*(\d+)", rest)
        kind = int(m2.group(1))
        rest = rest[m2.end():].lstrip()
    elif rest.startswith("("):
        end = _match_paren(rest, 0)
        inner = rest[1:end].replace(" ", "")
        rest = rest[end + 1:].lstrip()
        inner = inner.replace("kind=", "")
        if base != "character":
            kind = 8 if inner in ("double", "8", "dp", "kind(0.d0)", "kind(1.d0)") else (16 if inner == "16" else 4)
    if base == "integer":
        return "i", rest
    if base == "logical":
        return "l", rest
    if base == "character":
        return "s", rest
    if base == "doubleprecision":
        return "r8", rest
    if base == "real":
        return ("r8" if kind == 8 else "r4"), rest
    if base == "complex":
        return ("c8" if kind in (8, 16) else "c4"), rest


def _parse_dims(s):
    """'3,3' | ':,:' | 'n' | '*' | 'lo:hi' -> list of (lo_ast|None, hi_ast|None|'*'|':')"""
    dims = []
    for d in _split_top(s):
        d = d.strip()
        if d == ":":
            dims.append((None, ":"))
        elif d == "*":
            dims.append((None, "*"))
        else:
            parts = _split_top(d, ":")
            if len(parts) == 2:
                lo = parse_expr(parts[0]) if parts[0] else None
                hi = "*" if parts[1] == "*" else (parse_expr(parts[1]) if parts[1] else ":")
                dims.append((lo, hi))
            else:
                dims.append((None, parse_expr(d)))
    return dims


def parse_declaration(st):
    """returns list of Sym or None if `st` is not a type declaration"""
    if not _TYPE_RE.match(st):
        return None
    type_, rest = _parse_type_spec(st)
    if re.match(r"^(recursive\s+)?function\b", rest):
        return None
    attrs = {}
    if "::" in rest:
        left, ents = rest.split("::", 1)
        for a in _split_top(left):
            a = a.strip()
            if not a:
                continue
            if a.startswith("dimension"):
                attrs["dims"] = _parse_dims(a[a.index("(") + 1:_match_paren(a, a.index("("))])
            elif a.startswith("intent"):
                attrs["intent"] = a[a.index("(") + 1:-1].replace(" ", "")
            elif a in ("allocatable", "parameter"):
                attrs[a] = True
    else:
        ents = rest
    syms = []
    for e in _split_top(ents):
        e = e.strip()
        if not e:
            continue
        init = None
        m = re.match(r"^([a-z_]\w*)\s*(\(.*?\))?\s*(?:=(?!=)\s*(.*))?$", e)
        if not m:
            # name(dims)=init with nested parens
            mm = re.match(r"^([a-z_]\w*)\s*", e)
            name = mm.group(1)
            rest_e = e[mm.end():]
            dims = attrs.get("dims")
            if rest_e.startswith("("):
                end = _match_paren(rest_e, 0)
                dims = _parse_dims(rest_e[1:end])
                rest_e = rest_e[end + 1:].strip()
            if rest_e.startswith("="):
                init = parse_expr(rest_e[1:])
        else:
            name = m.group(1)
            dims = attrs.get("dims")
            if m.group(2):
                # make sure the parenthesis is balanced (regex is non-greedy)
                rest_e = e[len(name):].lstrip()
                end = _match_paren(rest_e, 0)
                dims = _parse_dims(rest_e[1:end])
                tail = rest_e[end + 1:].strip()
                init = parse_expr(tail[1:]) if tail.startswith("=") else None
            elif m.group(3) is not None:
                init = parse_expr(m.group(3))
        syms.append(Sym(name, type_, dims, attrs.get("allocatable", False), attrs.get("parameter", False), init,
                        attrs.get("intent")))
    return syms


_PROC_RE = re.compile(r"^(?:(recursive|pure|elemental)\s+)*(?:(integer|real|double\s*precision|complex|logical)\s*"
                      r"(?:\([^)]*\)|\*\s*\d+)?\s*)?(subroutine|function)\s+([a-z_]\w*)\s*(\(.*?\))?\s*(?:result\s*\(\s*(\w+)\s*\))?$")
_END_RE = re.compile(r"^end\s*(subroutine|function|module|program)?\b")


def _parse_use(st):
    m = re.match(r"^use\s+([a-z_]\w*)\s*(?:,\s*only\s*:\s*(.*))?$", st)
    if not m:
        raise Unsupported("use statement %r" % st)
    only = None
    if m.group(2) is not None:
        only = {}
        for it in _split_top(m.group(2)):
            it = it.strip()
            if not it:
                continue
            if "=>" in it:
                loc, rem = [x.strip() for x in it.split("=>")]
                only[loc] = rem
            else:
                only[it] = it
    return m.group(1), only


def parse_file(path):
    """-> list of Module (a `program` unit is returned as a module whose body statements are ignored)"""
    text = open(path, encoding="latin-1").read()
    stmts = logical_statements(text)
    mods, cur_mod, cur_proc, in_contains = [], None, None, False
    for no, st in stmts:
        if cur_mod is None:
            m = re.match(r"^(module|program)\s+([a-z_]\w*)$", st)
            if m:
                cur_mod, in_contains = Module(m.group(2), path), False
                cur_mod.is_program = m.group(1) == "program"
                mods.append(cur_mod)
            continue
        if cur_proc is None:
            if re.match(r"^end\s*(module|program)\b", st) or st == "end":
                cur_mod = None
                continue
            if st == "contains":
                in_contains = True
                continue
            if in_contains:
                m = _PROC_RE.match(st)
                if m:
                    args = [a.strip() for a in (m.group(5) or "()")[1:-1].split(",") if a.strip()]
                    kind = m.group(3)
                    result = m.group(6) or (m.group(4) if kind == "function" else None)
                    rtype = None
                    if m.group(2):
                        rtype = _parse_type_spec(st[st.index(m.group(2)):])[0]
                    cur_proc = Proc(m.group(4), kind, args, result, cur_mod, no, rtype)
                    cur_mod.procs[cur_proc.name] = cur_proc
                continue
            # module specification part
            if getattr(cur_mod, "is_program", False):
                continue
            if st.startswith("use "):
                cur_mod.uses.append(_parse_use(st))
                continue
            try:
                syms = parse_declaration(st)
            except Unsupported:
                syms = None
            if syms:
                for s in syms:
                    cur_mod.syms[s.name] = s
                    cur_mod.order.append(s.name)
            continue
        # inside a procedure
        m = _END_RE.match(st)
        if m and m.group(1) in ("subroutine", "function") or st == "end":
            cur_proc = None
            continue
        cur_proc.body_lines.append((no, st))
    return mods


# ------------------------------------------------------------------------------------------------------------
# statement parsing (procedure bodies)
# ------------------------------------------------------------------------------------------------------------

_SKIP_RE = re.compile(r"^(print\b|write\s*\(|format\s*\(|implicit\b|private\b|public\b|save\b|external\b|intrinsic\b|include\b|"
                      r"flush\b)")
_IO_RE = re.compile(r"^(open|close|read|rewind|backspace|inquire)\s*\(")


def _top_level_assign(st):
    """index of the top-level '=' of an assignment statement or -1"""
    depth, q = 0, None
    for i, ch in enumerate(st):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            prev = st[i - 1] if i else ""
            nxt = st[i + 1] if i + 1 < len(st) else ""
            if prev in "=<>/" or nxt in "=>":
                continue
            return i
    return -1


def parse_simple(st):
    """an action statement that can follow a one-line if"""
    if _SKIP_RE.match(st):
        return ("nop",)
    if _IO_RE.match(st):
        return ("unsupported", st)
    if st == "return":
        return ("return",)
    if st.startswith("stop"):
        return ("stop", st)
    if st == "cycle":
        return ("cycle",)
    if st == "exit":
        return ("exit",)
    if st == "continue":
        return ("nop",)
    m = re.match(r"^call\s+([a-z_]\w*)\s*(\(.*\))?$", st)
    if m:
        args = []
        if m.group(2):
            p = Parser(tokenize(m.group(2)))
            p.expect("(")
            args = p.arglist()
        return ("call", m.group(1), args)
    m = re.match(r"^(allocate|deallocate)\s*\((.*)\)$", st)
    if m:
        items = [parse_expr(x) for x in _split_top(m.group(2)) if not re.match(r"^\s*stat\s*=", x)]
        return (m.group(1), items)
    k = _top_level_assign(st)
    if k > 0:
        return ("assign", parse_expr(st[:k]), parse_expr(st[k + 1:]))
    return ("unsupported", st)


def parse_body(proc):
    """declarations into proc.syms / proc.uses, executable statements into a tree"""
    lines = proc.body_lines
    root, stack = [], []       # stack of (kind, node, current_block)
    block = root
    exec_started = False
    for no, st in lines:
        try:
            if not exec_started:
                if st.startswith("use "):
                    proc.uses.append(_parse_use(st))
                    continue
                if _SKIP_RE.match(st) and not st.startswith(("print", "write")):
                    continue
                syms = parse_declaration(st)
                if syms is not None:
                    for s in syms:
                        proc.syms[s.name] = s
                        proc.order.append(s.name)
                    continue
                exec_started = True
            # block constructs
            m = re.match(r"^if\s*\(", st)
            if m:
                end = _match_paren(st, st.index("("))
                cond = parse_expr(st[st.index("(") + 1:end])
                rest = st[end + 1:].strip()
                if rest == "then":
                    node = ["if", [(cond, [])], None]
                    block.append(node)
                    stack.append((node, block))
                    block = node[1][0][1]
                else:
                    block.append(["if", [(cond, [parse_simple(rest) + (no,)])], None])
                continue
            m = re.match(r"^else\s*if\s*\(", st)
            if m:
                end = _match_paren(st, st.index("("))
                cond = parse_expr(st[st.index("(") + 1:end])
                node = stack[-1][0]
                node[1].append((cond, []))
                block = node[1][-1][1]
                continue
            if st == "else":
                node = stack[-1][0]
                node[2] = []
                block = node[2]
                continue
            if re.match(r"^end\s*(if|do|select)$", st):
                node, block = stack.pop()
                continue
            m = re.match(r"^do\s+while\s*\((.*)\)$", st)
            if m:
                node = ["dowhile", parse_expr(m.group(1)), []]
                block.append(node)
                stack.append((node, block))
                block = node[2]
                continue
            if st == "do":
                node = ["dowhile", ("log", True), []]
                block.append(node)
                stack.append((node, block))
                block = node[2]
                continue
            m = re.match(r"^do\s+([a-z_]\w*)\s*=(.*)$", st)
            if m and len(_split_top(m.group(2))) >= 2:
                parts = _split_top(m.group(2))
                node = ["do", m.group(1), parse_expr(parts[0]), parse_expr(parts[1]),
                        parse_expr(parts[2]) if len(parts) > 2 else None, []]
                block.append(node)
                stack.append((node, block))
                block = node[5]
                continue
            m = re.match(r"^select\s*case\s*\((.*)\)$", st)
            if m:
                node = ["select", parse_expr(m.group(1)), []]
                block.append(node)
                stack.append((node, block))
                block = []           # statements before the first case are dropped
                continue
            m = re.match(r"^case\s*(default|\(.*\))$", st)
            if m:
                node = stack[-1][0]
                if m.group(1) == "default":
                    sel = None
                else:
                    p = Parser(tokenize(m.group(1)))
                    p.expect("(")
                    sel = p.arglist()
                node[2].append((sel, []))
                block = node[2][-1][1]
                continue
            block.append(parse_simple(st) + (no,))
        except Unsupported as e:
            block.append(("unsupported", "%s  [%s]" % (st, e), no))
    for a in proc.args:
        if a in proc.syms:
            proc.syms[a].is_dummy = True
    if proc.kind == "function" and proc.result not in proc.syms:
        proc.syms[proc.result] = Sym(proc.result, proc.rtype or "r4")
    proc.body = root


# ------------------------------------------------------------------------------------------------------------
# run-time support
# ------------------------------------------------------------------------------------------------------------

_DT = {"i": np.int64, "r4": np.float32, "r8": np.float64, "c4": np.complex64, "c8": np.complex128, "l": np.bool_, "s": object}
_F4, _F8, _C4, _C8 = np.float32, np.float64, np.complex64, np.complex128
_CPLX = (np.complex64, np.complex128, complex)


class FArray:
    """a Fortran array variable: `.a` is the column-major numpy storage (None while unallocated)"""
    __slots__ = ("a", "s")

    def __init__(self, a=None, s=None):
        self.a = a
        self.s = s          # rank-1 dummies: the whole storage sequence behind `a` (Fortran does not bounds-check)


def _zeros(shape, t):
    return FArray(np.zeros(shape, dtype=_DT[t], order="F"))


def _bind(arg, shape):
    """associate an actual array argument with an explicit-shape / assumed-size dummy (sequence association)"""
    if not isinstance(arg, FArray):
        arg = FArray(np.asfortranarray(arg))
    a = arg.a
    if a.shape == shape:
        if a.ndim == 1:
            arg.s = a if arg.s is None or arg.s.size < a.size or not np.shares_memory(arg.s, a) else arg.s
        return arg
    flat = a.reshape(-1, order="F")
    if a.size and not np.shares_memory(flat, a):
        raise Unsupported("sequence association of a non-contiguous section")
    if arg.s is not None and a.ndim == 1 and arg.s.size > flat.size:
        flat = arg.s
    if shape[-1] is None:
        lead = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
        shape = shape[:-1] + (flat.size // lead,)
    n = int(np.prod(shape))
    return FArray(flat[:n].reshape(shape, order="F"), flat if len(shape) == 1 else None)


def _seq(arr, *idx):
    """storage sequence starting at element arr(idx...) (for `call s(a(k))` with an array dummy)"""
    a = arr.a
    flat = a.reshape(-1, order="F")
    if a.size and not np.shares_memory(flat, a):
        raise Unsupported("sequence association of a non-contiguous array")
    off, mult = 0, 1
    for k, i in enumerate(idx):
        off += (i - 1) * mult
        mult *= a.shape[k]
    return FArray(flat[off:], flat[off:])


def _lin(arr, *idx):
    """element arr(idx...) addressed through the storage sequence WITHOUT bounds checks per dimension, as compiled
    Fortran does (geometry.f90 update_sigma indexes g_sigma(6,npt) as g_sigma(i<=npt, j<=6)); -> (value, setter)"""
    a = arr.a
    flat = a.reshape(-1, order="F")
    off, mult = 0, 1
    for k, i in enumerate(idx):
        off += (i - 1) * mult
        mult *= a.shape[k]

    def setter(v):
        flat[off] = v
    return flat[off], setter


def _wrap(v):
    if isinstance(v, np.ndarray):
        return FArray(v)
    return v


def _toi(x):
    if type(x) is int:
        return x
    if isinstance(x, _CPLX):
        x = x.real
    return int(x)


def _tor4(x):
    if isinstance(x, _CPLX):
        x = x.real
    return _F4(x)


def _tor8(x):
    if type(x) is _F8:
        return x
    if isinstance(x, _CPLX):
        x = x.real
    return _F8(x)


def _toc4(x):
    return _C4(x)


def _toc8(x):
    if type(x) is _C8:
        return x
    return _C8(x)


def _tol(x):
    return bool(x)


def _tos(x):
    return x


def _re_if_c(x):
    """value stored into a real/integer array: Fortran takes the real part of a complex right-hand side"""
    if isinstance(x, _CPLX):
        return x.real
    if isinstance(x, np.ndarray) and np.iscomplexobj(x):
        return x.real
    return x


def _cdiv(a, b):
    """complex division as gfortran emits it (-fcx-fortran-rules: Smith's algorithm, tree-complex.c
    expand_complex_div_wide); numpy's own scalar division multiplies by a reciprocal and can differ in the last bit"""
    single = not (isinstance(a, (_C8, _F8, complex)) or isinstance(b, (_C8, _F8, complex)))
    ft, ct = (_F4, _C4) if single else (_F8, _C8)
    ar, ai, br, bi = ft(a.real), ft(a.imag), ft(b.real), ft(b.imag)
    if abs(br) < abs(bi):
        ratio = br / bi
        div = (br * ratio) + bi
        tr = (ar * ratio) + ai
        ti = (ai * ratio) - ar
    else:
        ratio = bi / br
        div = (bi * ratio) + br
        tr = (ai * ratio) + ar
        ti = ai - (ar * ratio)
    return ct(complex(tr / div, ti / div))


def _div(a, b):
    if type(a) is int and type(b) is int:
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    if isinstance(b, _CPLX) and not isinstance(a, np.ndarray):
        return _cdiv(a, b)
    if isinstance(a, np.ndarray) and isinstance(b, (np.ndarray, int)) and a.dtype.kind == "i" and \
            (type(b) is int or b.dtype.kind == "i"):
        return np.trunc(a / b).astype(np.int64)
    return a / b


def _powi(x, n):
    """x**n for integer n as gfortran expands it (square-and-multiply, left to right)"""
    if n == 0:
        return x * 0 + 1
    if n < 0:
        return (x * 0 + 1) / _powi(x, -n) if type(x) is not int else (0 if abs(x) > 1 else int(x ** n))
    if n == 2:
        return x * x
    result = None
    base = x
    while n:
        if n & 1:
            result = base if result is None else result * base
        n >>= 1
        if n:
            base = base * base
    return result


def _pow(a, b):
    if type(b) is int:
        if type(a) is int:
            return a ** b if b >= 0 else _powi(a, b)
        return _powi(a, b)
    if type(a) is int:
        a = type(b)(a) if not isinstance(b, np.ndarray) else a
    if type(a) is _F8 and type(b) is _F8:
        return _F8(math.pow(a, b))            # libm pow, as gfortran calls it
    return np.power(a, b)


def _i_cmplx(x, y=None, kind=None):
    dt = _C8 if kind == 8 else _C4
    ft = _F8 if kind == 8 else _F4
    if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
        if y is None:
            return np.asarray(x).astype(dt)
        return (np.asarray(x).astype(ft).astype(dt) + dt(1j) * np.asarray(y).astype(ft).astype(dt)).astype(dt)
    if y is None:
        return dt(x)
    return dt(complex(ft(x), ft(y)))


def _i_real(x, kind=None):
    if isinstance(x, np.ndarray):
        if np.iscomplexobj(x):
            return x.real.copy() if kind is None else x.real.astype(_F8 if kind == 8 else _F4)
        return x.astype(_F8 if kind == 8 else _F4)
    if isinstance(x, _CPLX):
        r = x.real
        if kind is not None:
            return (_F8 if kind == 8 else _F4)(r)
        return r if not isinstance(x, complex) else _F8(r)
    return (_F8 if kind == 8 else _F4)(x)


def _i_dble(x):
    if isinstance(x, np.ndarray):
        return (x.real if np.iscomplexobj(x) else x).astype(_F8)
    return _F8(x.real if isinstance(x, _CPLX) else x)


def _i_aimag(x):
    return x.imag if not isinstance(x, np.ndarray) else x.imag.copy()


def _i_int(x, kind=None):
    if isinstance(x, np.ndarray):
        return np.trunc(x.real if np.iscomplexobj(x) else x).astype(np.int64)
    return _toi(x)


def _i_nint(x):
    if isinstance(x, np.ndarray):
        return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5)).astype(np.int64)
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _i_abs(x):
    if type(x) is int:
        return abs(x)
    return np.abs(x)


def _i_mod(a, p):
    if type(a) is int and type(p) is int:
        return int(math.fmod(a, p))
    return np.fmod(a, p)


def _i_sign(a, b):
    if type(a) is int and type(b) is int:
        return abs(a) if b >= 0 else -abs(a)
    return np.copysign(a, b)


def _i_min(*xs):
    r = xs[0]
    for x in xs[1:]:
        r = np.minimum(r, x) if isinstance(r, np.ndarray) or isinstance(x, np.ndarray) else (x if x < r else r)
    return r


def _i_max(*xs):
    r = xs[0]
    for x in xs[1:]:
        r = np.maximum(r, x) if isinstance(r, np.ndarray) or isinstance(x, np.ndarray) else (x if x > r else r)
    return r


def _scalar(v):
    return int(v) if isinstance(v, (np.integer,)) else v


def _i_maxval(x):
    return _scalar(np.max(x))


def _i_minval(x):
    return _scalar(np.min(x))


def _i_sum(x, dim=None):
    if dim is not None:
        return np.sum(x, axis=dim - 1)
    return _scalar(np.sum(x))


def _i_size(x, dim=None):
    x = x.a if isinstance(x, FArray) else x
    return int(x.size) if dim is None else int(x.shape[dim - 1])


def _i_allocated(x):
    return x.a is not None


def _i_sizeof(x):
    return 0


def _i_matmul(a, b):
    return np.matmul(a, b)


def _i_dot_product(a, b):
    return _scalar(np.sum(np.conj(a) * b) if np.iscomplexobj(a) else np.sum(a * b))


def _ac(items):
    flat = []
    for it in items:
        if isinstance(it, np.ndarray):
            flat.extend(it.reshape(-1, order="F").tolist() if it.dtype.kind == "i" else list(it.reshape(-1, order="F")))
        elif isinstance(it, list):
            flat.extend(it)
        else:
            flat.append(it)
    if all(type(v) is int for v in flat):
        return np.array(flat, dtype=np.int64)
    return np.array(flat)


def _frange(lo, hi, step=1):
    if step > 0:
        return range(lo, hi + 1, step)
    return range(lo, hi - 1, step)


def _after_loop(lo, hi, step=1):
    n = (hi - lo + step) // step
    return lo + max(n, 0) * step


def _assign_alloc(arr, rhs):
    """whole-array assignment to an allocatable (Fortran 2003: reallocates when the shape differs)"""
    if isinstance(rhs, np.ndarray):
        if arr.a is None or arr.a.shape != rhs.shape:
            dt = arr.a.dtype if arr.a is not None else rhs.dtype
            arr.a = np.zeros(rhs.shape, dtype=dt, order="F")
        arr.a[...] = rhs if not (arr.a.dtype.kind in "fi" and np.iscomplexobj(rhs)) else rhs.real
    else:
        arr.a[...] = _re_if_c(rhs) if arr.a.dtype.kind in "fi" else rhs


def _stop(msg=""):
    raise FortranStop(msg)


def _unsupported(msg):
    raise Unsupported(msg)


_INTRINSICS = {
    "abs": "_i_abs", "dabs": "_i_abs", "cabs": "_i_abs", "sqrt": "_i_sqrt", "dsqrt": "_i_sqrt", "csqrt": "np.sqrt",
    "sin": "_i_sin", "dsin": "_i_sin", "cos": "_i_cos", "dcos": "_i_cos", "tan": "_i_tan", "atan": "_i_atan",
    "datan": "_i_atan", "atan2": "_i_atan2", "asin": "np.arcsin", "acos": "np.arccos", "exp": "_i_exp", "dexp": "_i_exp",
    "cexp": "np.exp", "log": "_i_log", "dlog": "_i_log", "log10": "_i_log10", "sinh": "np.sinh", "cosh": "np.cosh",
    "tanh": "np.tanh", "conjg": "np.conj", "aimag": "_i_aimag", "dimag": "_i_aimag", "real": "_i_real", "dble": "_i_dble", "dfloat": "_i_dble",
    "cmplx": "_i_cmplx", "dcmplx": "_i_dcmplx", "int": "_i_int", "nint": "_i_nint", "mod": "_i_mod", "sign": "_i_sign",
    "min": "_i_min", "max": "_i_max", "dmin1": "_i_min", "dmax1": "_i_max", "maxval": "_i_maxval", "minval": "_i_minval",
    "sum": "_i_sum", "size": "_i_size", "allocated": "_i_allocated", "sizeof": "_i_sizeof", "matmul": "_i_matmul",
    "dot_product": "_i_dot_product", "transpose": "np.transpose", "floor": "_i_floor", "float": "_i_real",
}


def _libm(fmath, fnp):
    def f(x, *rest):
        t = type(x)
        if t is _F8 and all(type(r) in (_F8, int) for r in rest):
            return _F8(fmath(x, *rest))
        if t is _F4 and not rest:
            return _F4(fmath(float(x)))
        return fnp(x, *rest)
    return f


_i_sqrt = _libm(math.sqrt, np.sqrt)
_i_sin = _libm(math.sin, np.sin)
_i_cos = _libm(math.cos, np.cos)
_i_tan = _libm(math.tan, np.tan)
_i_atan = _libm(math.atan, np.arctan)
_i_atan2 = _libm(math.atan2, np.arctan2)
_i_exp = _libm(math.exp, np.exp)
_i_log = _libm(math.log, np.log)
_i_log10 = _libm(math.log10, np.log10)


def _i_dcmplx(x, y=None):
    return _i_cmplx(x, y, kind=8)


def _i_floor(x):
    return int(math.floor(x))


# ------------------------------------------------------------------------------------------------------------
# code generation
# ------------------------------------------------------------------------------------------------------------

_CONV = {"i": "_toi", "r4": "_tor4", "r8": "_tor8", "c4": "_toc4", "c8": "_toc8", "l": "_tol", "s": "_tos"}
_ZERO = {"i": "0", "r4": "_Z4", "r8": "_Z8", "c4": "_ZC4", "c8": "_ZC8", "l": "False", "s": "''"}


class ModNS:
    """attribute namespace holding a module's variables"""
    pass


class Runtime:
    def __init__(self, paths, skip_calls=(), hookable=(), linear=()):
        self.modules, self.ns = {}, {}
        self.skip_calls = set(skip_calls)
        self.linear = set(linear)       # arrays addressed through their storage sequence (out-of-bounds subscripts)
        self.hookable, self.hooks = set(hookable), {}
        self.consts = []
        self.env = dict(np=np, math=math, FArray=FArray, _FA=FArray, _zeros=_zeros, _bind=_bind, _seq=_seq, _wrap=_wrap, _lin=_lin,
                        _toi=_toi, _tor4=_tor4, _tor8=_tor8, _toc4=_toc4, _toc8=_toc8, _tol=_tol, _tos=_tos, _re_if_c=_re_if_c,
                        _div=_div, _pow=_pow, _ac=_ac, _frange=_frange, _after_loop=_after_loop, _assign_alloc=_assign_alloc,
                        _stop=_stop, _unsupported=_unsupported, _Z4=_F4(0), _Z8=_F8(0), _ZC4=_C4(0), _ZC8=_C8(0), _int=int,
                        _K=self.consts, _rt=self)
        for k, v in list(globals().items()):
            if k.startswith("_i_"):
                self.env[k] = v
        for p in paths:
            for m in parse_file(p):
                self.modules[m.name] = m
        for m in self.modules.values():
            self.ns[m.name] = ModNS()
            self.env["m_" + m.name] = self.ns[m.name]
        self.call_counts = {}
        done = set()
        for m in list(self.modules.values()):
            self._init_module(m, done)

    # ---- symbol resolution -------------------------------------------------------------------------------
    def resolve_in_module(self, mod, name, seen=None):
        """-> ('var', module, Sym) | ('proc', module, Proc) | None, following module-level use association"""
        seen = seen or set()
        if (mod.name, name) in seen:
            return None
        seen.add((mod.name, name))
        if name in mod.syms:
            return ("var", mod, mod.syms[name])
        if name in mod.procs:
            return ("proc", mod, mod.procs[name])
        for mname, only in mod.uses:
            um = self.modules.get(mname)
            if um is None:
                continue
            if only is None:
                r = self.resolve_in_module(um, name, seen)
            elif name in only:
                r = self.resolve_in_module(um, only[name], seen)
            else:
                r = None
            if r:
                return r
        return None

    def resolve(self, proc, mod, name):
        if proc is not None:
            if name in proc.syms:
                return ("local", None, proc.syms[name])
            for mname, only in proc.uses:
                um = self.modules.get(mname)
                if um is None:
                    continue
                if only is None:
                    r = self.resolve_in_module(um, name)
                elif name in only:
                    r = self.resolve_in_module(um, only[name])
                else:
                    r = None
                if r:
                    return r
        return self.resolve_in_module(mod, name)

    # ---- module initialisation ---------------------------------------------------------------------------
    def _init_module(self, m, done):
        if m.name in done:
            return
        done.add(m.name)
        for mname, _ in m.uses:
            if mname in self.modules:
                self._init_module(self.modules[mname], done)
        ns = self.ns[m.name]
        for name in m.order:
            s = m.syms[name]
            try:
                if s.dims:
                    if s.allocatable or any(d[1] in (":", "*") for d in s.dims):
                        setattr(ns, name, FArray(None))
                    else:
                        shape = tuple(self._eval_const(d[1], m) for d in s.dims)
                        arr = _zeros(shape, s.type)
                        if s.init is not None:
                            arr.a[...] = self._eval_const(s.init, m)
                        setattr(ns, name, arr)
                elif m.name == "kind_param" and name in ("single", "double", "short", "long"):
                    setattr(ns, name, {"single": 4, "double": 8, "short": 2, "long": 4}[name])
                elif s.init is not None:
                    setattr(ns, name, self.env[_CONV[s.type]](self._eval_const(s.init, m)))
                else:
                    setattr(ns, name, eval(_ZERO[s.type], self.env))
            except Exception as e:      # noqa: BLE001  (declarations outside the subset are simply left unset)
                setattr(ns, name, None)

    def _eval_const(self, ast, mod):
        g = Gen(self, None, mod)
        return eval(g.expr(ast), self.env)

    # ---- procedures --------------------------------------------------------------------------------------
    def get_proc(self, modname, name):
        r = self.resolve_in_module(self.modules[modname], name)
        if not r or r[0] != "proc":
            raise KeyError("%s::%s" % (modname, name))
        return r[2]

    def pyfunc(self, proc):
        if proc.pyfunc is None:
            if proc.body is None:
                parse_body(proc)
            src = Gen(self, proc, proc.module).function()
            proc.pysrc = src
            code = compile(src, "<f90:%s:%s>" % (proc.module.name, proc.name), "exec")
            loc = {}
            exec(code, self.env, loc)
            proc.pyfunc = loc["f_" + proc.name]
            self.env["f_%s__%s" % (proc.module.name, proc.name)] = proc.pyfunc
        return proc.pyfunc

    def signature(self, proc):
        if proc.body is None:
            parse_body(proc)
        return proc

    def call(self, modname, name, *args):
        """call a procedure from Python; array arguments are numpy arrays / FArray (shared), scalars by value.
        Returns the function result, or for a subroutine the tuple of its scalar out-values."""
        proc = self.get_proc(modname, name)
        f = self.pyfunc(proc)

        def conv(a):
            if isinstance(a, np.ndarray):
                return FArray(a)
            if type(a) is float:
                return _F8(a)
            if type(a) is complex:
                return _C8(a)
            return a
        return f(*[conv(a) for a in args])

    def mod(self, name):
        return self.ns[name]

    def run_hook(self, name):
        h = self.hooks.get(name)
        if h is not None:
            h()


class Gen:
    def __init__(self, rt, proc, mod):
        self.rt, self.proc, self.mod = rt, proc, mod
        self.lines, self.ind = [], 1
        self.tmp = 0
        self.ido_vars = []

    # ---- helpers ----
    def emit(self, s):
        self.lines.append("    " * self.ind + s)

    def const(self, value):
        self.rt.consts.append(value)
        return "_K[%d]" % (len(self.rt.consts) - 1)

    def lookup(self, name):
        if name in self.ido_vars:
            return ("ido", None, None)
        return self.rt.resolve(self.proc, self.mod, name)

    def varref(self, name):
        r = self.lookup(name)
        if r is None:
            raise Unsupported("unknown name %r in %s" % (name, self.proc.name if self.proc else self.mod.name))  # noqa
        if r[0] == "ido":
            return "i_" + name, None
        if r[0] == "local":
            return "v_" + name, r[2]
        if r[0] == "var":
            return "m_%s.%s" % (r[1].name, r[2].name), r[2]
        raise Unsupported("%r is a procedure, not a variable" % name)

    # ---- expressions ----
    def expr(self, e):
        k = e[0]
        if k == "int":
            return str(e[1])
        if k == "real":
            return self.const(_F8(e[1]) if e[2] == 8 else _F4(e[1]))
        if k == "log":
            return "True" if e[1] else "False"
        if k == "str":
            return repr(e[1])
        if k == "paren":
            return "(" + self.expr(e[1]) + ")"
        if k == "cplx":
            re_, im_ = e[1], e[2]
            dbl = any(self._is_double_literal(x) for x in (re_, im_))
            return "_i_cmplx(%s, %s, %s)" % (self.expr(re_), self.expr(im_), "8" if dbl else "None")
        if k == "name":
            ref, sym = self.varref(e[1])
            if sym is not None and sym.rank:
                return ref + ".a"
            return ref
        if k == "un":
            if e[1] == "not":
                return "(not %s)" % self.expr(e[2])
            return "(%s%s)" % (e[1], self.expr(e[2]))
        if k == "bin":
            op, a, b = e[1], self.expr(e[2]), self.expr(e[3])
            if op == "/":
                return "_div(%s, %s)" % (a, b)
            if op == "**":
                return "_pow(%s, %s)" % (a, b)
            if op in ("and", "or"):
                return "(%s %s %s)" % (a, op, b)
            if op == ".eqv.":
                return "(bool(%s) == bool(%s))" % (a, b)
            if op == ".neqv.":
                return "(bool(%s) != bool(%s))" % (a, b)
            if op == "//":
                return "(%s + %s)" % (a, b)
            return "(%s %s %s)" % (a, op, b)
        if k == "ac":
            return "_ac([%s])" % ", ".join(self.ac_item(x) for x in e[1])
        if k == "call":
            return self.call_or_index(e)
        if k == "range":
            raise Unsupported("array section outside a subscript")
        raise Unsupported("expression node %r" % (k,))

    def _is_double_literal(self, e):
        if e[0] == "real":
            return e[2] == 8
        if e[0] == "un":
            return self._is_double_literal(e[2])
        return False

    def ac_item(self, it):
        if it[0] == "ido":
            _, items, var, lo, hi, step = it
            lo_, hi_ = self.expr(lo), self.expr(hi)
            st_ = self.expr(step) if step else "1"
            self.ido_vars.append(var)
            inner = ", ".join(self.ac_item(x) for x in items)
            self.ido_vars.pop()
            if len(items) == 1:
                return "[%s for i_%s in _frange(%s, %s, %s)]" % (inner, var, lo_, hi_, st_)
            return "[_x for i_%s in _frange(%s, %s, %s) for _x in (%s,)]" % (var, lo_, hi_, st_, inner)
        return self.expr(it)

    def subscript(self, sym, args):
        """numpy index expression for sym(args); returns (index_text, is_element)"""
        parts, element = [], True
        for k, a in enumerate(args):
            lb = None
            if sym.dims and k < len(sym.dims) and sym.dims[k][0] is not None:
                lb = self.expr(sym.dims[k][0])
            off = "1" if lb is None else "(%s)" % lb
            if a[0] == "range":
                element = False
                lo = "" if a[1] is None else "(%s)-%s" % (self.expr(a[1]), off)
                hi = "" if a[2] is None else ("(%s)-%s+1" % (self.expr(a[2]), off) if off != "1" else self.expr(a[2]))
                st = "" if a[3] is None else ":" + self.expr(a[3])
                parts.append("%s:%s%s" % (lo, hi, st))
            elif self.rank_of(a) > 0:
                element = False
                parts.append("(%s)-%s" % (self.expr(a), off))
            else:
                ex = self.expr(a)
                parts.append("%s-%s" % (ex if re.match(r"^[\w.]+$", ex) else "(" + ex + ")", off))
        return ", ".join(parts), element

    def rank_of(self, e):
        """static rank (0 = scalar); only what subscripts / actual arguments need"""
        k = e[0]
        if k == "name":
            r = self.lookup(e[1])
            if r and r[0] in ("local", "var") and r[2] is not None:
                return r[2].rank
            return 0
        if k == "ac":
            return 1
        if k == "paren":
            return self.rank_of(e[1])
        if k == "un":
            return self.rank_of(e[2])
        if k == "bin":
            return max(self.rank_of(e[2]), self.rank_of(e[3]))
        if k == "call":
            r = self.lookup(e[1])
            if r and r[0] in ("local", "var") and r[2].rank:
                return sum(1 for a in e[2] if a[0] == "range" or self.rank_of(a) > 0)
            if r and r[0] == "proc":
                p = self.rt.signature(r[2])
                s = p.syms.get(p.result)
                return s.rank if s else 0
            if e[1] in ("maxval", "minval", "sum", "size", "dot_product", "allocated", "sizeof"):
                return 0
            if e[1] in _INTRINSICS and e[2]:
                return max(self.rank_of(a) for a in e[2] if a[0] != "kw")
        return 0

    @staticmethod
    def storage(sym, element):
        """rank-1 explicit-shape dummies are addressed through the whole storage sequence for single elements:
        the reference reads past the declared extent of such dummies (merge_sort's a(n1)), which Fortran allows
        silently"""
        if element and sym.is_dummy and sym.rank == 1 and sym.dims[0][1] not in (":",):
            return "s"
        return "a"

    def call_or_index(self, e):
        name, args = e[1], e[2]
        r = self.lookup(name)
        if r is not None and r[0] in ("local", "var") and r[2].type != "proc":
            ref, sym = self.varref(name)
            if not sym.rank:
                raise Unsupported("subscripted scalar %r" % name)
            idx, element = self.subscript(sym, args)
            if element and sym.name in self.rt.linear:
                return "_lin(%s, %s)[0]" % (ref, ", ".join(self.expr(x) for x in args))
            st = self.storage(sym, element)
            if element and sym.type == "i":
                return "_int(%s.%s[%s])" % (ref, st, idx)
            return "%s.%s[%s]" % (ref, st, idx)
        if r is not None and r[0] == "proc":
            return self.proc_call(r[2], args, as_function=True)
        if name in _INTRINSICS:
            return self.intrinsic(name, args)
        raise Unsupported("unknown function or array %r" % name)

    def intrinsic(self, name, args):
        pos, kw = [], {}
        for a in args:
            if a[0] == "kw":
                kw[a[1]] = a[2]
            else:
                pos.append(a)
        if name in ("allocated", "size", "sizeof"):
            ref, _ = self.varref(pos[0][1]) if pos[0][0] == "name" else (None, None)
            if ref is None:
                return "0" if name == "sizeof" else "_i_size(%s)" % self.expr(pos[0])
            extra = "".join(", " + self.expr(a) for a in pos[1:])
            return "%s(%s%s)" % (_INTRINSICS[name], ref, extra)
        if name in ("real", "cmplx", "int"):
            kind = kw.get("kind")
            n_data = 2 if name == "cmplx" else 1
            if kind is None and len(pos) > n_data and name != "cmplx":
                kind = pos.pop()
            elif kind is None and name == "cmplx" and len(pos) == 3:
                kind = pos.pop()
            kind_s = "None"
            if kind is not None:
                kind_s = "8" if (kind == ("name", "double") or kind == ("int", 8)) else "4"
            if name == "cmplx":
                a0 = self.expr(pos[0])
                a1 = self.expr(pos[1]) if len(pos) > 1 else "None"
                return "_i_cmplx(%s, %s, %s)" % (a0, a1, kind_s)
            return "%s(%s, %s)" % (_INTRINSICS[name], self.expr(pos[0]), kind_s)
        return "%s(%s)" % (_INTRINSICS[name], ", ".join(self.expr(a) for a in pos))

    def actual(self, a, dummy):
        """text of one actual argument; dummy = Sym of the callee's dummy (or None)"""
        if a[0] == "name":
            r = self.lookup(a[1])
            if r and r[0] in ("local", "var") and r[2] is not None and r[2].rank:
                return self.varref(a[1])[0]               # the FArray object itself
            if r and r[0] == "proc":
                raise Unsupported("procedure as argument")
            return self.expr(a)
        if a[0] == "call":
            r = self.lookup(a[1])
            if r and r[0] in ("local", "var") and r[2].rank:
                ref, sym = self.varref(a[1])
                idx, element = self.subscript(sym, a[2])
                if element and dummy is not None and dummy.rank:
                    return "_seq(%s, %s)" % (ref, ", ".join(self.expr(x) for x in a[2]))
                if not element:
                    return "_FA(%s.a[%s])" % (ref, idx)
        return "_wrap(%s)" % self.expr(a)

    def proc_call(self, callee, args, as_function):
        sig = self.rt.signature(callee)
        actuals = []
        for k, a in enumerate(args):
            if a[0] == "kw":
                raise Unsupported("keyword arguments in a procedure call")
            dummy = sig.syms.get(sig.args[k]) if k < len(sig.args) else None
            actuals.append(self.actual(a, dummy))
        key = "f_%s__%s" % (callee.module.name, callee.name)
        # late binding through a cached global: first call translates the callee
        return "%s(%s)" % (self._callee_ref(callee, key), ", ".join(actuals))

    def _callee_ref(self, callee, key):
        if key not in self.rt.env:
            rt = self.rt

            def trampoline(*a, _callee=callee, _key=key):
                f = rt.pyfunc(_callee)
                return f(*a)
            self.rt.env[key] = trampoline
        return key

    # ---- statements ----
    def assign(self, lhs, rhs):
        if lhs[0] == "name":
            ref, sym = self.varref(lhs[1])
            val = self.expr(rhs)
            if sym is None:
                self.emit("%s = %s" % (ref, val))
            elif sym.rank:
                if sym.allocatable:
                    self.emit("_assign_alloc(%s, %s)" % (ref, val))
                elif sym.type in ("i", "r4", "r8"):
                    self.emit("%s.a[...] = _re_if_c(%s)" % (ref, val))
                else:
                    self.emit("%s.a[...] = %s" % (ref, val))
            else:
                self.emit("%s = %s(%s)" % (ref, _CONV[sym.type], val))
            return
        if lhs[0] == "call":
            ref, sym = self.varref(lhs[1])
            if not sym.rank:
                raise Unsupported("assignment to %r" % (lhs[1],))
            idx, element = self.subscript(sym, lhs[2])
            val = self.expr(rhs)
            if element and sym.name in self.rt.linear:
                self.emit("_lin(%s, %s)[1](%s)" % (ref, ", ".join(self.expr(x) for x in lhs[2]), val))
                return
            if sym.type in ("i", "r4", "r8"):
                val = "_re_if_c(%s)" % val
            self.emit("%s.%s[%s] = %s" % (ref, self.storage(sym, element), idx, val))
            return
        raise Unsupported("assignment target %r" % (lhs,))

    def call_stmt(self, name, args):
        if name in self.rt.skip_calls:
            self.emit("pass")
            return
        if name in self.rt.hookable:
            self.emit("_rt.run_hook(%r)" % name)      # test tap: runs just before the call
        r = self.lookup(name)
        if r is None or r[0] != "proc":
            raise Unsupported("call of unknown subroutine %r" % name)
        callee = r[2]
        sig = self.rt.signature(callee)
        call = self.proc_call(callee, args, as_function=False)
        outs = []
        for k in sig.out_scalars():
            if k < len(args) and args[k][0] in ("name", "call"):
                a = args[k]
                rr = self.lookup(a[1])
                if rr and rr[0] in ("local", "var") and rr[2] is not None:
                    if a[0] == "name" and not rr[2].rank and not rr[2].parameter:
                        outs.append((k, a))
                        continue
                    if a[0] == "call" and rr[2].rank:
                        idx, element = self.subscript(rr[2], a[2])
                        d = sig.syms.get(sig.args[k])
                        if element and not (d is not None and d.rank):
                            outs.append((k, a))
                            continue
            outs.append((k, None))
        if not any(a for _, a in outs):
            self.emit(call)
            return
        self.tmp += 1
        t = "_r%d" % self.tmp
        self.emit("%s = %s" % (t, call))
        for pos, (k, a) in enumerate(outs):
            if a is None:
                continue
            if a[0] == "name":
                ref, sym = self.varref(a[1])
                self.emit("%s = %s[%d]" % (ref, t, pos))
            else:
                ref, sym = self.varref(a[1])
                idx, _ = self.subscript(sym, a[2])
                self.emit("%s.a[%s] = %s[%d]" % (ref, idx, t, pos))

    def ret(self):
        p = self.proc
        if p.kind == "function":
            s = p.syms[p.result]
            self.emit("return v_%s%s" % (p.result, ".a" if s.rank else ""))
        else:
            outs = p.out_scalars()
            if outs:
                self.emit("return (%s,)" % ", ".join("v_" + p.args[k] for k in outs))
            else:
                self.emit("return None")

    def block(self, stmts):
        if not stmts:
            self.emit("pass")
            return
        for st in stmts:
            self.stmt(st)

    def stmt(self, st):
        k = st[0]
        try:
            if k == "nop":
                self.emit("pass")
            elif k == "assign":
                self.assign(st[1], st[2])
            elif k == "call":
                self.call_stmt(st[1], st[2])
            elif k == "return":
                self.ret()
            elif k == "stop":
                self.emit("_stop(%r)" % st[1])
            elif k == "cycle":
                self.emit("continue")
            elif k == "exit":
                self.emit("break")
            elif k == "unsupported":
                self.emit("_unsupported(%r)" % (st[1],))
            elif k == "allocate":
                for it in st[1]:
                    ref, sym = self.varref(it[1])
                    dims = []
                    for a in it[2]:
                        if a[0] == "range":
                            raise Unsupported("allocate with explicit lower bound")
                        dims.append(self.expr(a))
                    self.emit("%s.a = np.zeros((%s,), dtype=np.%s, order='F')" % (ref, ", ".join(dims), _DT[sym.type].__name__))
            elif k == "deallocate":
                for it in st[1]:
                    ref, _ = self.varref(it[1])
                    self.emit("%s.a = None" % ref)
            elif k == "if":
                for n, (cond, blk) in enumerate(st[1]):
                    self.emit("%s %s:" % ("if" if n == 0 else "elif", self.expr(cond)))
                    self.ind += 1
                    self.block(blk)
                    self.ind -= 1
                if st[2] is not None:
                    self.emit("else:")
                    self.ind += 1
                    self.block(st[2])
                    self.ind -= 1
            elif k == "do":
                _, var, lo, hi, step, blk = st
                ref, sym = self.varref(var)
                self.tmp += 1
                t = self.tmp
                self.emit("_lo%d = %s; _hi%d = %s; _st%d = %s" % (t, self.expr(lo), t, self.expr(hi), t, self.expr(step) if step else "1"))
                self.emit("for %s in _frange(_lo%d, _hi%d, _st%d):" % (ref, t, t, t))
                self.ind += 1
                self.block(blk)
                self.ind -= 1
                self.emit("else:")
                self.emit("    %s = _after_loop(_lo%d, _hi%d, _st%d)" % (ref, t, t, t))
            elif k == "dowhile":
                self.emit("while %s:" % self.expr(st[1]))
                self.ind += 1
                self.block(st[2])
                self.ind -= 1
            elif k == "select":
                self.tmp += 1
                t = "_s%d" % self.tmp
                self.emit("%s = %s" % (t, self.expr(st[1])))
                first = True
                default = None
                for sel, blk in st[2]:
                    if sel is None:
                        default = blk
                        continue
                    conds = []
                    for a in sel:
                        if a[0] == "range":
                            c = []
                            if a[1] is not None:
                                c.append("%s >= %s" % (t, self.expr(a[1])))
                            if a[2] is not None:
                                c.append("%s <= %s" % (t, self.expr(a[2])))
                            conds.append("(" + " and ".join(c) + ")")
                        else:
                            conds.append("%s == %s" % (t, self.expr(a)))
                    self.emit("%s %s:" % ("if" if first else "elif", " or ".join(conds)))
                    first = False
                    self.ind += 1
                    self.block(blk)
                    self.ind -= 1
                if default is not None:
                    if first:
                        self.block(default)
                    else:
                        self.emit("else:")
                        self.ind += 1
                        self.block(default)
                        self.ind -= 1
            else:
                raise Unsupported("statement %r" % (k,))
        except Unsupported as e:
            self.emit("_unsupported(%r)" % ("line %s: %s" % (st[-1] if isinstance(st[-1], int) else "?", e),))

    def function(self):
        p = self.proc
        self.lines = ["def f_%s(%s):" % (p.name, ", ".join("a_" + a for a in p.args))]
        # dummies
        for a in p.args:
            s = p.syms.get(a)
            if s is None or not s.rank:
                self.emit("v_%s = a_%s" % (a, a))
        for a in p.args:
            s = p.syms.get(a)
            if s is not None and s.rank:
                if all(d[1] == ":" for d in s.dims):
                    self.emit("v_%s = a_%s" % (a, a))
                else:
                    shape = ", ".join("None" if d[1] == "*" else self.expr(d[1]) for d in s.dims)
                    self.emit("v_%s = _bind(a_%s, (%s,))" % (a, a, shape))
        # locals
        for name in p.order:
            s = p.syms[name]
            if s.is_dummy:
                continue
            if s.rank:
                if s.allocatable or any(d[1] in (":", "*") for d in s.dims):
                    self.emit("v_%s = _FA(None)" % name)
                else:
                    shape = ", ".join(self.expr(d[1]) for d in s.dims)
                    self.emit("v_%s = _zeros((%s,), %r)" % (name, shape, s.type))
                    if s.init is not None:
                        self.emit("v_%s.a[...] = %s" % (name, self.expr(s.init)))
            elif s.init is not None:
                self.emit("v_%s = %s(%s)" % (name, _CONV[s.type], self.expr(s.init)))
            else:
                self.emit("v_%s = %s" % (name, _ZERO[s.type]))
        if p.kind == "function" and p.result not in p.order:
            self.emit("v_%s = %s" % (p.result, _ZERO[p.syms[p.result].type]))
        self.block(p.body)
        self.ret()
        return "\n".join(self.lines) + "\n"
