"""
`ExternalField`: an analytic field added to (or replacing) one gathered field component of the particles at
every step, right after the gather (fbpic/lpa_utils/external_fields.py:13-215, call site main.py:472-473).

The reference compiles the user's Python function to a GPU kernel with Numba.  There is no Numba on this path:
the function body is translated to CUDA C by a small AST walk (arithmetic, comparisons, conditional
expressions, `math.*` / `np.*` functions and constants, local assignments, if/else with returns) and compiled by
NVRTC inside libfbpic_b200.so (`b2_external_field_compile`), or the user passes the CUDA C expression as a
string.  A function outside that subset is rejected with an explanatory error -- it is never evaluated on the
host.
"""
import ast
import ctypes
import inspect
import math
import textwrap
import numpy as np
from scipy.constants import c

from .. import _lib
from .._lib import call

ARGS = ('F', 'x', 'y', 'z', 't', 'amplitude', 'length_scale')

# python name -> CUDA C function (double precision overloads of the CUDA math library)
_FUNCS = {'sin': 'sin', 'cos': 'cos', 'tan': 'tan', 'asin': 'asin', 'acos': 'acos', 'atan': 'atan',
          'arcsin': 'asin', 'arccos': 'acos', 'arctan': 'atan', 'atan2': 'atan2', 'arctan2': 'atan2',
          'sinh': 'sinh', 'cosh': 'cosh', 'tanh': 'tanh', 'exp': 'exp', 'log': 'log', 'log10': 'log10',
          'sqrt': 'sqrt', 'fabs': 'fabs', 'abs': 'fabs', 'absolute': 'fabs', 'floor': 'floor', 'ceil': 'ceil',
          'pow': 'pow', 'power': 'pow', 'hypot': 'hypot', 'erf': 'erf', 'erfc': 'erfc', 'fmod': 'fmod',
          'copysign': 'copysign', 'sign': '_b2_sign', 'minimum': 'fmin', 'maximum': 'fmax', 'min': 'fmin',
          'max': 'fmax', 'expm1': 'expm1', 'log1p': 'log1p', 'cbrt': 'cbrt'}
_CONSTS = {'pi': math.pi, 'e': math.e, 'inf': float('inf')}
_BINOPS = {ast.Add: '+', ast.Sub: '-', ast.Mult: '*', ast.Div: '/'}
_CMPOPS = {ast.Lt: '<', ast.LtE: '<=', ast.Gt: '>', ast.GtE: '>=', ast.Eq: '==', ast.NotEq: '!='}


class TranslationError(ValueError):
    pass


class _Translator(object):
    """Python function (AST) -> CUDA C statements over the scalars F, x, y, z, t, amplitude, length_scale."""

    def __init__(self, func):
        self.func = func
        try:
            src = textwrap.dedent(inspect.getsource(func))
        except (OSError, TypeError):
            raise TranslationError('the source of `field_func` is not available: pass the CUDA C expression as '
                                   'a string instead')
        tree = ast.parse(src)
        fdefs = [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.Lambda))]
        if not fdefs:
            raise TranslationError('no function definition found in the source of `field_func`')
        self.fdef = fdefs[0]
        params = [a.arg for a in self.fdef.args.args]
        if len(params) != 7:
            raise TranslationError('`field_func` must take the 7 arguments (F, x, y, z, t, amplitude, length_scale)')
        self.names = dict(zip(params, ARGS))      # positional: the user's own argument names are free
        self.locals = set()
        closure = inspect.getclosurevars(func) if inspect.isfunction(func) else None
        self.outer = {}
        if closure is not None:
            self.outer.update(closure.globals)
            self.outer.update(closure.nonlocals)

    def fail(self, node, what):
        raise TranslationError('`field_func`: %s (line %d) is outside the subset that is translated to CUDA C '
                               '(arithmetic, comparisons, `a if c else b`, math./np. functions, local '
                               'assignments, if/else); pass a CUDA C expression string instead'
                               % (what, getattr(node, 'lineno', 0)))

    # ---- expressions
    def expr(self, n):
        if isinstance(n, ast.Constant):
            if isinstance(n.value, bool):
                return '1' if n.value else '0'
            if isinstance(n.value, (int, float)):
                return self.number(n.value)
            self.fail(n, 'constant %r' % (n.value,))
        if isinstance(n, ast.Name):
            if n.id in self.names:
                return self.names[n.id]
            if n.id in self.locals:
                return 'v_' + n.id
            if n.id in self.outer and isinstance(self.outer[n.id], (int, float, np.floating, np.integer)):
                return self.number(float(self.outer[n.id]))        # captured numeric constant
            self.fail(n, 'name `%s`' % n.id)
        if isinstance(n, ast.Attribute):
            # `mod.name`: evaluated on the object the function really captured (scipy.constants.e is the elementary
            # charge, math.e is Euler's number); the bare spelling math.pi / np.pi of an un-captured module keeps working
            if isinstance(n.value, ast.Name):
                base = self.outer.get(n.value.id)
                if base is not None:
                    val = getattr(base, n.attr, None)
                    if isinstance(val, (int, float, np.floating, np.integer)) and not isinstance(val, bool):
                        return self.number(float(val))
                    self.fail(n, 'attribute `%s.%s`' % (n.value.id, n.attr))
                if n.value.id in ('math', 'np', 'numpy') and n.attr in _CONSTS:
                    return self.number(_CONSTS[n.attr])
            self.fail(n, 'attribute `%s`' % n.attr)
        if isinstance(n, ast.UnaryOp):
            if isinstance(n.op, ast.USub):
                return '(-%s)' % self.expr(n.operand)
            if isinstance(n.op, ast.UAdd):
                return self.expr(n.operand)
            if isinstance(n.op, ast.Not):
                return '(!%s)' % self.expr(n.operand)
            self.fail(n, 'unary operator')
        if isinstance(n, ast.BinOp):
            a, b = self.expr(n.left), self.expr(n.right)
            if type(n.op) in _BINOPS:
                return '(%s %s %s)' % (a, _BINOPS[type(n.op)], b)
            if isinstance(n.op, ast.Pow):
                if isinstance(n.right, ast.Constant) and n.right.value == 2:
                    return '(%s * %s)' % (a, a)
                return 'pow(%s, %s)' % (a, b)
            if isinstance(n.op, ast.Mod):
                return '_b2_pymod(%s, %s)' % (a, b)
            self.fail(n, 'binary operator')
        if isinstance(n, ast.Compare):
            parts, left = [], n.left
            for op, right in zip(n.ops, n.comparators):
                if type(op) not in _CMPOPS:
                    self.fail(n, 'comparison operator')
                parts.append('(%s %s %s)' % (self.expr(left), _CMPOPS[type(op)], self.expr(right)))
                left = right
            return '(' + ' && '.join(parts) + ')'
        if isinstance(n, ast.BoolOp):
            op = ' && ' if isinstance(n.op, ast.And) else ' || '
            return '(' + op.join(self.expr(v) for v in n.values) + ')'
        if isinstance(n, ast.IfExp):
            return '(%s ? %s : %s)' % (self.expr(n.test), self.expr(n.body), self.expr(n.orelse))
        if isinstance(n, ast.Call):
            f = n.func
            name = f.attr if isinstance(f, ast.Attribute) else (f.id if isinstance(f, ast.Name) else None)
            if name in ('float', 'float64') and len(n.args) == 1:
                return self.expr(n.args[0])
            if name not in _FUNCS or n.keywords:
                self.fail(n, 'call of `%s`' % name)
            return '%s(%s)' % (_FUNCS[name], ', '.join('(double)' + self.expr(a) for a in n.args))
        self.fail(n, type(n).__name__)

    @staticmethod
    def number(v):
        v = float(v)
        if math.isinf(v):
            return '(1.0/0.0)' if v > 0 else '(-1.0/0.0)'
        return repr(v) if ('e' in repr(v) or '.' in repr(v)) else repr(v) + '.0'

    # ---- statements
    def block(self, stmts, indent):
        out = []
        for st in stmts:
            if isinstance(st, ast.Expr) and isinstance(st.value, ast.Constant) and isinstance(st.value.value, str):
                continue                                  # docstring
            if isinstance(st, ast.Return):
                if st.value is None:
                    self.fail(st, 'empty return')
                out.append('%sF_[i_] = %s; return;' % (indent, self.expr(st.value)))
            elif isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                tgt = st.targets[0].id
                val = self.expr(st.value)
                if tgt in self.names:
                    self.fail(st, 'assignment to the argument `%s`' % tgt)
                if tgt in self.locals:
                    out.append('%sv_%s = %s;' % (indent, tgt, val))
                else:
                    self.locals.add(tgt)
                    self.decls.append(tgt)
                    out.append('%sv_%s = %s;' % (indent, tgt, val))
            elif isinstance(st, ast.AugAssign) and isinstance(st.target, ast.Name) and type(st.op) in _BINOPS \
                    and st.target.id in self.locals:
                out.append('%sv_%s %s= %s;' % (indent, st.target.id, _BINOPS[type(st.op)], self.expr(st.value)))
            elif isinstance(st, ast.If):
                out.append('%sif (%s) {' % (indent, self.expr(st.test)))
                out += self.block(st.body, indent + '    ')
                if st.orelse:
                    out.append('%s} else {' % indent)
                    out += self.block(st.orelse, indent + '    ')
                out.append('%s}' % indent)
            elif isinstance(st, ast.Pass):
                continue
            else:
                self.fail(st, type(st).__name__ + ' statement')
        return out

    def translate(self):
        self.decls = []
        if isinstance(self.fdef, ast.Lambda):
            body = ['    F_[i_] = %s; return;' % self.expr(self.fdef.body)]
        else:
            body = self.block(self.fdef.body, '    ')
        decl = ['    double %s;' % ', '.join('v_' + d for d in self.decls)] if self.decls else []
        return '\n'.join(decl + body)


_HELPERS = ('    auto _b2_sign = [](double v) { return (double)((v > 0.) - (v < 0.)); };\n'
            '    auto _b2_pymod = [](double a, double b) { double r = fmod(a, b); return (r != 0. && ((r < 0.) != (b < 0.))) ? r + b : r; };\n'
            '    (void)_b2_sign; (void)_b2_pymod;\n')


def python_to_cuda(field_func):
    """CUDA C statements (over the scalars F, x, y, z, t, amplitude, length_scale; result stored to
    `F_[i_]`) equivalent to the Python function or to the CUDA C expression string `field_func`."""
    if isinstance(field_func, str):
        return _HELPERS + '    F_[i_] = %s;' % field_func
    return _HELPERS + _Translator(field_func).translate()


class ExternalField(object):

    def __init__(self, field_func, fieldtype, amplitude, length_scale, species=None, gamma_boost=None):
        """`field_func(F, x, y, z, t, amplitude, length_scale)` returns the modified field F' (lab frame);
        use `return F + ...` to add to the gathered field.  Either a Python function written with `math`
        functions (as the reference requires for its GPU path, external_fields.py:50-53) or a CUDA C
        expression string.  Same arguments and boosted-frame behaviour as the reference (:13-181)."""
        self.length_scale = length_scale
        self.species = species
        if fieldtype not in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
            raise ValueError("`fieldtype` must be one of Ex, Ey, Ez, Bx, By, Bz")
        self.cuda_source = python_to_cuda(field_func)
        self.gamma_boost, self.beta_boost = 1., 0.
        inv_c = 1. / c
        if (gamma_boost is not None) and (gamma_boost != 1.):
            # the expression is evaluated at the lab-frame (z, t) of the particle and the resulting lab-frame
            # field is Lorentz-transformed through its amplitude (external_fields.py:118-181)
            g = gamma_boost
            self.gamma_boost, self.beta_boost = g, np.sqrt(1. - 1. / g**2)
            gb = g * self.beta_boost
            pairs = {'Ex': (('Ex', g * amplitude), ('By', -gb * inv_c * amplitude)),
                     'Ey': (('Ey', g * amplitude), ('Bx', gb * inv_c * amplitude)),
                     'Bx': (('Bx', g * amplitude), ('Ey', gb * c * amplitude)),
                     'By': (('By', g * amplitude), ('Ex', -gb * c * amplitude)),
                     'Ez': (('Ez', amplitude),), 'Bz': (('Bz', amplitude),)}
            self.fieldtypes_and_amplitudes = pairs[fieldtype]
        else:
            self.fieldtypes_and_amplitudes = ((fieldtype, amplitude),)
        self._handle = None

    def _kernel(self):
        if self._handle is None:
            h = ctypes.c_void_p()
            call.b2_external_field_compile(self.cuda_source.encode(), ctypes.byref(h))
            self._handle = h
        return self._handle

    def apply_expression(self, ptcl, t):
        """Apply the expression to the gathered field of the particles (external_fields.py:183-215)."""
        for species in ptcl:
            if (self.species is None) or (species is self.species):
                if species.Ntot <= 0:
                    continue
                species._need_gpu()
                for fieldtype, amplitude in self.fieldtypes_and_amplitudes:
                    field = getattr(species, fieldtype)
                    call.b2_external_field_apply(_lib.context().handle, self._kernel(), species.Ntot, field.ptr,
                                                 species.x.ptr, species.y.ptr, species.z.ptr, t, amplitude,
                                                 self.length_scale, self.gamma_boost, self.beta_boost, None)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().b2_external_field_free(self._handle)
        except Exception:
            pass
