"""Run functions out of a CPython-3.7 .pyc (TEST INFRASTRUCTURE, used by make_reference_vectors.py only).

The reference's no-CBF module `MPC_optimize_kin` exists only as
PKG/__pycache__/MPC_optimize_kin.cpython-37.pyc; this interpreter (3.12) can neither import it nor
unmarshal its code objects.  This file reads the 3.7 marshal stream into plain structures and
interprets the small subset of 3.7 bytecode those methods use, so that the reference's own
`__init__`, `initialize_constraints` and `optimize_problem` can be executed (on the sympy-backed
casadi stand-in) instead of being restated from a disassembly.
"""
from __future__ import annotations

import operator
import struct


class Code:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return f"<code37 {self.name} line {self.firstlineno}>"


class _Reader:
    """marshal version 4 as written by CPython 3.7"""

    def __init__(self, data):
        self.b, self.i, self.refs = data, 0, []

    def u8(self):
        v = self.b[self.i]; self.i += 1; return v

    def i32(self):
        v = struct.unpack_from("<i", self.b, self.i)[0]; self.i += 4; return v

    def raw(self, n):
        v = self.b[self.i:self.i + n]; self.i += n; return v

    def obj(self):
        t = self.u8()
        flag, t = t & 0x80, chr(t & 0x7F)
        idx = None
        if flag:
            idx = len(self.refs)
            self.refs.append(None)
        v = self._load(t)
        if idx is not None:
            self.refs[idx] = v
        return v

    def _load(self, t):
        if t == "0": return None  # TYPE_NULL
        if t == "N": return None
        if t == "T": return True
        if t == "F": return False
        if t == ".": return Ellipsis
        if t == "i": return self.i32()
        if t == "g": v = struct.unpack_from("<d", self.b, self.i)[0]; self.i += 8; return v
        if t == "l":
            n = self.i32(); sign = -1 if n < 0 else 1; v = 0
            for k in range(abs(n)):
                d = struct.unpack_from("<H", self.b, self.i)[0]; self.i += 2; v |= d << (15 * k)
            return sign * v
        if t == "s": return bytes(self.raw(self.i32()))
        if t in "tu": return self.raw(self.i32()).decode("utf-8", "surrogatepass")
        if t in "aA": return self.raw(self.i32()).decode("latin-1")
        if t in "zZ": return self.raw(self.u8()).decode("latin-1")
        if t == ")": return tuple(self.obj() for _ in range(self.u8()))
        if t == "(": return tuple(self.obj() for _ in range(self.i32()))
        if t == "[": return [self.obj() for _ in range(self.i32())]
        if t == "r": return self.refs[self.i32()]
        if t == "c":
            argcount, kwonly, nlocals, stacksize, flags = (self.i32() for _ in range(5))
            code, consts, names, varnames, freevars, cellvars = (self.obj() for _ in range(6))
            filename, name = self.obj(), self.obj()
            firstlineno, lnotab = self.i32(), self.obj()
            return Code(argcount=argcount, kwonlyargcount=kwonly, nlocals=nlocals, flags=flags, code=code, consts=consts, names=names,
                        varnames=varnames, freevars=freevars, cellvars=cellvars, filename=filename, name=name, firstlineno=firstlineno,
                        lnotab=lnotab)
        raise ValueError(f"marshal type {t!r} at {self.i}")


def load_pyc(path) -> Code:
    data = open(path, "rb").read()
    assert data[:4] == b"\x42\x0d\x0d\x0a", "not a CPython 3.7 pyc"
    return _Reader(data[16:]).obj()


def find_code(code: Code, *path) -> Code:
    """nested code object by names, e.g. find_code(module, 'MPC_optimize', 'optimize_problem')"""
    for name in path:
        code = next(c for c in code.consts if isinstance(c, Code) and c.name == name)
    return code


# ---- the 3.7 opcodes the reference's methods use (numbers from CPython 3.7's opcode.py)
OP = {1: "POP_TOP", 2: "ROT_TWO", 3: "ROT_THREE", 4: "DUP_TOP", 10: "UNARY_POSITIVE", 11: "UNARY_NEGATIVE", 12: "UNARY_NOT",
      19: "BINARY_POWER", 20: "BINARY_MULTIPLY", 22: "BINARY_MODULO", 23: "BINARY_ADD", 24: "BINARY_SUBTRACT", 25: "BINARY_SUBSCR",
      26: "BINARY_FLOOR_DIVIDE", 27: "BINARY_TRUE_DIVIDE", 55: "INPLACE_ADD", 56: "INPLACE_SUBTRACT", 57: "INPLACE_MULTIPLY",
      29: "INPLACE_TRUE_DIVIDE", 60: "STORE_SUBSCR", 68: "GET_ITER", 80: "BREAK_LOOP", 83: "RETURN_VALUE", 87: "POP_BLOCK",
      92: "UNPACK_SEQUENCE", 93: "FOR_ITER", 95: "STORE_ATTR", 100: "LOAD_CONST", 102: "BUILD_TUPLE", 103: "BUILD_LIST", 105: "BUILD_MAP",
      106: "LOAD_ATTR", 107: "COMPARE_OP", 110: "JUMP_FORWARD", 111: "JUMP_IF_FALSE_OR_POP", 112: "JUMP_IF_TRUE_OR_POP",
      113: "JUMP_ABSOLUTE", 114: "POP_JUMP_IF_FALSE", 115: "POP_JUMP_IF_TRUE", 116: "LOAD_GLOBAL", 119: "CONTINUE_LOOP", 120: "SETUP_LOOP",
      124: "LOAD_FAST", 125: "STORE_FAST", 131: "CALL_FUNCTION", 133: "BUILD_SLICE", 141: "CALL_FUNCTION_KW", 142: "CALL_FUNCTION_EX",
      144: "EXTENDED_ARG",
      145: "LIST_APPEND", 156: "BUILD_CONST_KEY_MAP", 160: "LOAD_METHOD", 161: "CALL_METHOD"}
BIN = {"BINARY_POWER": operator.pow, "BINARY_MULTIPLY": operator.mul, "BINARY_MODULO": operator.mod, "BINARY_ADD": operator.add,
       "BINARY_SUBTRACT": operator.sub, "BINARY_SUBSCR": operator.getitem, "BINARY_FLOOR_DIVIDE": operator.floordiv,
       "BINARY_TRUE_DIVIDE": operator.truediv, "INPLACE_ADD": operator.iadd, "INPLACE_SUBTRACT": operator.isub,
       "INPLACE_MULTIPLY": operator.imul, "INPLACE_TRUE_DIVIDE": operator.itruediv}
CMP = [operator.lt, operator.le, operator.eq, operator.ne, operator.gt, operator.ge, lambda a, b: a in b, lambda a, b: a not in b,
       operator.is_, operator.is_not]


def disassemble(code: Code):
    out, ext, i = [], 0, 0
    while i < len(code.code):
        op, arg = code.code[i], code.code[i + 1] | ext
        ext = (arg << 8) if op == 144 else 0
        out.append((i, OP.get(op, f"<{op}>"), arg))
        i += 2
    return out


def run(code: Code, glb: dict, *args, **kwargs):
    """execute a function code object with positional/keyword arguments (no defaults, closures or generators)"""
    import builtins

    loc = [None] * code.nlocals
    for k, a in enumerate(args):
        loc[k] = a
    for k, v in kwargs.items():
        loc[code.varnames.index(k)] = v
    stack, blocks, pc, ext = [], [], 0, 0
    bc = code.code
    while True:
        op, arg = bc[pc], bc[pc + 1] | ext
        name = OP.get(op)
        if name is None:
            raise NotImplementedError(f"{code.name}: opcode {op} at {pc}")
        nxt, ext = pc + 2, 0
        if name == "EXTENDED_ARG":
            ext = arg << 8
        elif name == "LOAD_CONST":
            stack.append(code.consts[arg])
        elif name == "LOAD_FAST":
            stack.append(loc[arg])
        elif name == "STORE_FAST":
            loc[arg] = stack.pop()
        elif name == "LOAD_GLOBAL":
            n = code.names[arg]
            stack.append(glb[n] if n in glb else getattr(builtins, n))
        elif name in ("LOAD_ATTR", "LOAD_METHOD"):
            stack.append(getattr(stack.pop(), code.names[arg]))
        elif name == "STORE_ATTR":
            obj = stack.pop(); setattr(obj, code.names[arg], stack.pop())
        elif name == "STORE_SUBSCR":
            k = stack.pop(); obj = stack.pop(); obj[k] = stack.pop()
        elif name in BIN:
            b = stack.pop(); a = stack.pop(); stack.append(BIN[name](a, b))
        elif name == "UNARY_NEGATIVE":
            stack.append(-stack.pop())
        elif name == "UNARY_POSITIVE":
            stack.append(+stack.pop())
        elif name == "UNARY_NOT":
            stack.append(not stack.pop())
        elif name == "COMPARE_OP":
            b = stack.pop(); a = stack.pop(); stack.append(CMP[arg](a, b))
        elif name in ("CALL_FUNCTION", "CALL_METHOD"):
            a = [stack.pop() for _ in range(arg)][::-1]
            stack.append(stack.pop()(*a))
        elif name == "CALL_FUNCTION_KW":
            keys = stack.pop()
            a = [stack.pop() for _ in range(arg)][::-1]
            npos = arg - len(keys)
            stack.append(stack.pop()(*a[:npos], **dict(zip(keys, a[npos:]))))
        elif name == "CALL_FUNCTION_EX":
            kw = stack.pop() if arg & 1 else {}
            a = stack.pop()
            stack.append(stack.pop()(*a, **kw))
        elif name == "BUILD_TUPLE":
            v = tuple(stack[len(stack) - arg:]); del stack[len(stack) - arg:]; stack.append(v)
        elif name == "BUILD_LIST":
            v = list(stack[len(stack) - arg:]); del stack[len(stack) - arg:]; stack.append(v)
        elif name == "BUILD_MAP":
            items = stack[len(stack) - 2 * arg:]; del stack[len(stack) - 2 * arg:]
            stack.append({items[2 * k]: items[2 * k + 1] for k in range(arg)})
        elif name == "BUILD_CONST_KEY_MAP":
            keys = stack.pop(); vals = stack[len(stack) - arg:]; del stack[len(stack) - arg:]
            stack.append(dict(zip(keys, vals)))
        elif name == "BUILD_SLICE":
            v = stack[len(stack) - arg:]; del stack[len(stack) - arg:]; stack.append(slice(*v))
        elif name == "UNPACK_SEQUENCE":
            stack.extend(list(stack.pop())[::-1])
        elif name == "LIST_APPEND":
            v = stack.pop(); stack[-arg].append(v)
        elif name == "POP_TOP":
            stack.pop()
        elif name == "DUP_TOP":
            stack.append(stack[-1])
        elif name == "ROT_TWO":
            stack[-1], stack[-2] = stack[-2], stack[-1]
        elif name == "ROT_THREE":
            stack[-1], stack[-2], stack[-3] = stack[-2], stack[-3], stack[-1]
        elif name == "GET_ITER":
            stack.append(iter(stack.pop()))
        elif name == "FOR_ITER":
            try:
                stack.append(next(stack[-1]))
            except StopIteration:
                stack.pop(); nxt = pc + 2 + arg
        elif name == "SETUP_LOOP":
            blocks.append((pc + 2 + arg, len(stack)))
        elif name == "POP_BLOCK":
            blocks.pop()
        elif name == "BREAK_LOOP":
            nxt, depth = blocks.pop(); del stack[depth:]
        elif name in ("JUMP_ABSOLUTE", "CONTINUE_LOOP"):
            nxt = arg
        elif name == "JUMP_FORWARD":
            nxt = pc + 2 + arg
        elif name == "POP_JUMP_IF_FALSE":
            if not stack.pop(): nxt = arg
        elif name == "POP_JUMP_IF_TRUE":
            if stack.pop(): nxt = arg
        elif name == "JUMP_IF_FALSE_OR_POP":
            if not stack[-1]: nxt = arg
            else: stack.pop()
        elif name == "JUMP_IF_TRUE_OR_POP":
            if stack[-1]: nxt = arg
            else: stack.pop()
        elif name == "RETURN_VALUE":
            return stack.pop()
        pc = nxt
