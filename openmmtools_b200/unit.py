"""A small dimensional-analysis layer with the part of ``openmm.unit``'s surface the replica-exchange path uses.

The reference's objects are unit bearing (``ThermodynamicState(system, temperature=300*unit.kelvin)``,
``LangevinSplittingDynamicsMove(timestep=1.0*unit.femtosecond)``, ``SamplerState.positions`` is a
``Quantity`` in nanometers; /root/reference/openmmtools/states.py:2022-2036, mcmc.py:1280).  OpenMM is not a
dependency of this package, so this module provides ``Quantity``/``Unit`` with the same spelling
(``value_in_unit``, ``value_in_unit_system``, ``unit.md_unit_system``, ``is_quantity`` ...).  Objects from a
real ``openmm.unit`` are accepted anywhere through :func:`to_md`.

The md unit system is OpenMM's: nm, ps, dalton (= g/mol), K, mol, e  ->  energies in kJ/mol.
"""
import math
import numpy as np

_DIMS = ('length', 'time', 'mass', 'temperature', 'amount', 'charge')


class Unit:
    __slots__ = ('dims', 'scale', 'name')

    def __init__(self, dims, scale, name):
        self.dims = tuple(dims)
        self.scale = float(scale)   # value of 1 <this unit> expressed in md base units
        self.name = name

    # -- algebra
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit([a + b for a, b in zip(self.dims, other.dims)], self.scale * other.scale,
                        _join(self.name, other.name, '*'))
        if isinstance(other, Quantity):
            return Quantity(other._value, self * other.unit)
        return Quantity(other, self)

    __rmul__ = lambda self, other: Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit([a - b for a, b in zip(self.dims, other.dims)], self.scale / other.scale,
                        _join(self.name, other.name, '/'))
        if isinstance(other, Quantity):
            return Quantity(1.0 / other._value, self / other.unit)
        return Quantity(1.0 / other, self)

    def __rtruediv__(self, other):
        return Quantity(other, self ** -1)

    def __pow__(self, p):
        return Unit([a * p for a in self.dims], self.scale ** p, '%s**%g' % (self.name, p) if self.name else '')

    def is_compatible(self, other):
        return all(abs(a - b) < 1e-12 for a, b in zip(self.dims, other.dims))

    def is_dimensionless(self):
        return all(abs(a) < 1e-12 for a in self.dims)

    def conversion_factor_to(self, other):
        if not self.is_compatible(other):
            raise TypeError('Unit "%s" is not compatible with Unit "%s".' % (self.name, other.name))
        return self.scale / other.scale

    def __eq__(self, other):
        return isinstance(other, Unit) and self.is_compatible(other) and math.isclose(self.scale, other.scale, rel_tol=1e-12)

    def __hash__(self):
        return hash((tuple(round(d, 9) for d in self.dims), float('%.12g' % self.scale)))

    def __repr__(self):
        return 'Unit(%s)' % (self.name or 'dimensionless')

    __str__ = lambda self: self.name or 'dimensionless'

    def get_name(self):
        return self.name


def _join(a, b, op):
    if not a:
        return b if op == '*' else '/' + b
    if not b:
        return a
    return '%s%s%s' % (a, op, b)


def _base(i, scale, name):
    d = [0] * 6
    d[i] = 1
    return Unit(d, scale, name)


dimensionless = Unit([0] * 6, 1.0, '')
nanometer = nanometers = _base(0, 1.0, 'nanometer')
angstrom = angstroms = _base(0, 0.1, 'angstrom')
meter = meters = _base(0, 1e9, 'meter')
centimeter = centimeters = _base(0, 1e7, 'centimeter')
picosecond = picoseconds = _base(1, 1.0, 'picosecond')
femtosecond = femtoseconds = _base(1, 1e-3, 'femtosecond')
nanosecond = nanoseconds = _base(1, 1e3, 'nanosecond')
second = seconds = _base(1, 1e12, 'second')
day = days = _base(1, 86400e12, 'day')
gram = grams = _base(2, 1.0, 'gram')
kilogram = kilograms = _base(2, 1e3, 'kilogram')
kelvin = kelvins = _base(3, 1.0, 'kelvin')
mole = moles = _base(4, 1.0, 'mole')
elementary_charge = elementary_charges = _base(5, 1.0, 'elementary charge')
item = Unit(mole.dims, 1.0 / 6.02214076e23, 'item')
dalton = daltons = amu = amus = Unit((gram / mole).dims, 1.0, 'dalton')
joule = joules = Unit((kilogram * meter ** 2 / second ** 2).dims, 1e-3, 'joule')
kilojoule = kilojoules = Unit(joule.dims, 1.0, 'kilojoule')
kilocalorie = kilocalories = Unit(joule.dims, 4.184, 'kilocalorie')
kilojoule_per_mole = kilojoules_per_mole = Unit((kilojoule / mole).dims, 1.0, 'kilojoule/mole')
kilocalorie_per_mole = kilocalories_per_mole = Unit((kilojoule / mole).dims, 4.184, 'kilocalorie/mole')
bar = bars = Unit((joule / meter ** 3).dims, 1e5 * 1e-3 / 1e27, 'bar')
atmosphere = atmospheres = Unit(bar.dims, 1.01325 * bar.scale, 'atmosphere')
radian = radians = Unit([0] * 6, 1.0, 'radian')
degree = degrees = Unit([0] * 6, math.pi / 180.0, 'degree')

AVOGADRO_CONSTANT_NA = None   # set below (needs Quantity)
BOLTZMANN_CONSTANT_kB = None
MOLAR_GAS_CONSTANT_R = None


class UnitSystem:
    def __init__(self, name):
        self.name = name

    def express_unit(self, u):
        return Unit(u.dims, 1.0, 'md(%s)' % u.name)


md_unit_system = UnitSystem('md')


class Quantity:
    """value * unit; value may be a float, a list or a numpy array."""
    __array_priority__ = 100

    def __init__(self, value=None, unit=None):
        if unit is None:
            if isinstance(value, Quantity):
                value, unit = value._value, value.unit
            elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], Quantity):
                unit = value[0].unit
                value = [q.value_in_unit(unit) for q in value]
            else:
                unit = dimensionless
        elif isinstance(value, Quantity):
            unit = value.unit * unit
            value = value._value
        elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], Quantity):
            inner = value[0].unit
            value = [_rows_value(q, inner) for q in value]
            unit = inner * unit
        self._value = value
        self.unit = unit

    # -- conversions
    def value_in_unit(self, unit):
        f = self.unit.conversion_factor_to(unit)
        return _scale(self._value, f)

    def value_in_unit_system(self, system):
        return _scale(self._value, self.unit.scale)

    def in_units_of(self, unit):
        return Quantity(self.value_in_unit(unit), unit)

    def in_unit_system(self, system):
        return Quantity(self.value_in_unit_system(system), system.express_unit(self.unit))

    def _md(self):
        return _scale(self._value, self.unit.scale)

    # -- arithmetic
    def __add__(self, other):
        other = _as_q(other)
        return Quantity(_arr(self._value) + other.value_in_unit(self.unit), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        other = _as_q(other)
        return Quantity(_arr(self._value) - other.value_in_unit(self.unit), self.unit)

    def __rsub__(self, other):
        other = _as_q(other)
        return Quantity(other.value_in_unit(self.unit) - _arr(self._value), self.unit)

    def __neg__(self):
        return Quantity(-_arr(self._value), self.unit)

    def __abs__(self):
        return Quantity(abs(_arr(self._value)), self.unit)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self._value, self.unit * other)
        if isinstance(other, Quantity):
            return _reduce(_arr(self._value) * _arr(other._value), self.unit * other.unit)
        return Quantity(_arr(self._value) * other, self.unit)

    def __rmul__(self, other):
        return Quantity(other * _arr(self._value), self.unit)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return _reduce(self._value, self.unit / other)
        if isinstance(other, Quantity):
            return _reduce(_arr(self._value) / _arr(other._value), self.unit / other.unit)
        return Quantity(_arr(self._value) / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / _arr(self._value), self.unit ** -1)

    def __pow__(self, p):
        return Quantity(_arr(self._value) ** p, self.unit ** p)

    def sqrt(self):
        return Quantity(np.sqrt(_arr(self._value)), self.unit ** 0.5)

    # -- comparisons
    def _cmp(self, other):
        other = _as_q(other)
        return _arr(self._value), other.value_in_unit(self.unit)

    def __eq__(self, other):
        if other is None or (not isinstance(other, Quantity) and not _is_number(other)):
            return False
        o = _as_q(other)
        if not self.unit.is_compatible(o.unit):
            return False
        a, b = self._cmp(o)
        r = a == b
        return bool(np.all(r)) if isinstance(r, np.ndarray) and r.ndim == 0 else r

    def __ne__(self, other):
        r = self.__eq__(other)
        return ~r if isinstance(r, np.ndarray) else not r

    def __lt__(self, other):
        a, b = self._cmp(other); return a < b

    def __le__(self, other):
        a, b = self._cmp(other); return a <= b

    def __gt__(self, other):
        a, b = self._cmp(other); return a > b

    def __ge__(self, other):
        a, b = self._cmp(other); return a >= b

    __hash__ = None

    # -- container behaviour
    def __len__(self):
        return len(self._value)

    def __getitem__(self, key):
        return Quantity(self._value[key], self.unit)

    def __setitem__(self, key, value):
        if isinstance(value, Quantity):
            value = value.value_in_unit(self.unit)
        self._value[key] = value

    def __iter__(self):
        for v in self._value:
            yield Quantity(v, self.unit)

    def __float__(self):
        if not self.unit.is_dimensionless():
            raise TypeError('only dimensionless quantities convert to float')
        return float(self._value) * self.unit.scale

    def __bool__(self):
        return bool(np.any(self._value))

    def __copy__(self):
        return Quantity(self._value, self.unit)

    def __deepcopy__(self, memo):
        import copy
        return Quantity(copy.deepcopy(self._value, memo), self.unit)

    def __repr__(self):
        return 'Quantity(value=%r, unit=%s)' % (self._value, self.unit)

    def __str__(self):
        return '%s %s' % (self._value, self.unit)

    def __getstate__(self):
        return {'_value': self._value, 'dims': self.unit.dims, 'scale': self.unit.scale, 'name': self.unit.name}

    def __setstate__(self, s):
        self._value = s['_value']
        self.unit = Unit(s['dims'], s['scale'], s['name'])

    @property
    def shape(self):
        return np.shape(self._value)


def _rows_value(q, unit):
    return q.value_in_unit(unit) if isinstance(q, Quantity) else q


def _is_number(x):
    return isinstance(x, (int, float, np.number, np.ndarray, list, tuple))


def _arr(v):
    return np.asarray(v) if isinstance(v, (list, tuple)) else v


def _scale(v, f):
    if f == 1.0:
        return np.array(v) if isinstance(v, (list, tuple)) else v
    return _arr(v) * f


def _as_q(x):
    return x if isinstance(x, Quantity) else Quantity(x, dimensionless)


def _reduce(value, unit):
    """A dimensionless result collapses to a plain number (as openmm.unit does for e.g. kT/kT)."""
    if unit.is_dimensionless():
        return _scale(value, unit.scale)
    return Quantity(value, unit)


def is_quantity(x):
    return isinstance(x, Quantity) or (hasattr(x, 'value_in_unit_system') and hasattr(x, 'unit'))


def is_unit(x):
    return isinstance(x, Unit)


def sqrt(x):
    return x.sqrt() if isinstance(x, Quantity) else math.sqrt(x)


def to_md(x, unit=None, name='value'):
    """Plain md-unit number/array from a Quantity of this module, an ``openmm.unit`` Quantity, or a bare number.

    ``unit`` (a Unit of this module) is checked for dimensional compatibility when ``x`` is one of our
    Quantities; bare numbers are taken to be in md units already.
    """
    if x is None:
        return None
    if isinstance(x, Quantity):
        if unit is not None and not x.unit.is_compatible(unit):
            raise TypeError('%s must have units compatible with %s, got %s' % (name, unit, x.unit))
        return x._md()
    if hasattr(x, 'value_in_unit_system'):          # openmm.unit.Quantity
        try:
            import openmm.unit as ou
            return x.value_in_unit_system(ou.md_unit_system)
        except ImportError:                          # pragma: no cover
            raise TypeError('foreign quantity without openmm installed')
    return np.asarray(x, dtype=np.float64) if isinstance(x, (list, tuple, np.ndarray)) else float(x)


AVOGADRO_CONSTANT_NA = Quantity(6.02214076e23, mole ** -1)
BOLTZMANN_CONSTANT_kB = Quantity(1.380649e-23, joule / kelvin)
MOLAR_GAS_CONSTANT_R = Quantity(8.31446261815324e-3, kilojoule_per_mole / kelvin)
