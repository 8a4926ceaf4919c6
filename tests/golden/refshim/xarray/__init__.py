"""Minimal stand-in for the parts of xarray that lmadaus/efa_xray touches.

TEST INFRASTRUCTURE ONLY.  xarray is not installable in this image (no network,
not in the wheelhouse), so tests/golden/make_golden.py puts this directory on
sys.path to import and run the UNMODIFIED reference package from /root/reference
and record its outputs as golden vectors.  Only documented xarray semantics that
the reference relies on ("xarray surface actually used")
are implemented; anything else raises.  Never imported by the product package.
"""
import copy as _copy
from collections import OrderedDict
import numpy as np


class Variable(object):
    """dims + ndarray; plays the role of both xarray.Variable and DataArray."""

    def __init__(self, dims, values):
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims)
        self._values = np.asarray(values)
        assert self._values.ndim == len(self.dims), (self.dims, self._values.shape)

    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, new):
        new = np.asarray(new)
        assert new.shape == self._values.shape, (new.shape, self._values.shape)
        self._values = new

    @property
    def shape(self):
        return self._values.shape

    def __len__(self):
        return self._values.shape[0]

    def __getitem__(self, key):
        out = self._values[key]
        # integer indexing drops dims, slices keep them
        if not isinstance(key, tuple):
            key = (key,)
        dims = []
        k = 0
        for d in self.dims:
            if k < len(key):
                # slices and integer ARRAYS keep the dimension (xarray: da[int array] is a DataArray over the same dim)
                if isinstance(key[k], slice) or np.ndim(key[k]) == 1:
                    dims.append(d)
                k += 1
            else:
                dims.append(d)
        return Variable(tuple(dims), out)

    def __setitem__(self, key, val):
        if isinstance(val, Variable):
            val = val.values
        self._values[key] = val

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._values, dtype=dtype)

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims if d in self.dims]
        return Variable(tuple(self.dims[i] for i in order), self._values.transpose(order))

    def mean(self, dim=None, axis=None):
        if dim is not None:
            ax = self.dims.index(dim)
            return Variable(tuple(d for d in self.dims if d != dim), self._values.mean(axis=ax))
        return Variable((), self._values.mean(axis=axis)) if axis is None else self._values.mean(axis=axis)

    def _binary(self, other, op):
        if isinstance(other, Variable):
            # broadcast by dimension name (other's dims must be a subset, in order)
            idx = tuple(slice(None) if d in other.dims else None for d in self.dims)
            assert [d for d in self.dims if d in other.dims] == list(other.dims)
            return Variable(self.dims, op(self._values, other.values[idx]))
        return Variable(self.dims, op(self._values, other))

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def to_dataset(self, dim='variable'):
        assert self.dims[0] == dim
        names = self._names
        ds = Dataset()
        for i, n in enumerate(names):
            ds._data_vars[n] = Variable(self.dims[1:], self._values[i])
        ds._coords = self._coords
        return ds


def DataArray(values, coords=None):
    """DataArray(v, [(dimname, coordvalues)]) as used at assimilation.py:90."""
    dims = tuple(c[0] for c in coords)
    return Variable(dims, values)


class _Mapping(object):
    def __init__(self, *dicts):
        self._dicts = dicts

    def __getitem__(self, k):
        for d in self._dicts:
            if k in d:
                return d[k]
        raise KeyError(k)

    def __contains__(self, k):
        return any(k in d for d in self._dicts)

    def keys(self):
        out = []
        for d in self._dicts:
            out.extend(d.keys())
        return out

    def items(self):
        return [(k, self[k]) for k in self.keys()]


class Dataset(object):
    def __init__(self, data_vars=None, coords=None):
        self._data_vars = OrderedDict()
        self._coords = OrderedDict()
        for k, v in (coords or {}).items():
            if isinstance(v, tuple):
                self._coords[k] = Variable(v[0], v[1])
            else:
                self._coords[k] = Variable((k,), v)
        for k, v in (data_vars or {}).items():
            self._data_vars[k] = Variable(v[0], v[1])

    # --- mapping views
    @property
    def coords(self):
        return _Mapping(self._coords)

    @property
    def variables(self):
        # xarray lists coordinate variables together with data variables
        return _Mapping(self._data_vars, self._coords)

    def keys(self):
        return self.variables.keys()

    def __getitem__(self, k):
        return self.variables[k]

    def __deepcopy__(self, memo):
        new = Dataset()
        new.__class__ = self.__class__
        new._data_vars = OrderedDict((k, Variable(v.dims, v.values.copy())) for k, v in self._data_vars.items())
        new._coords = OrderedDict((k, Variable(v.dims, v.values.copy())) for k, v in self._coords.items())
        return new

    # --- methods the reference calls
    def to_array(self):
        names = list(self._data_vars.keys())
        first = self._data_vars[names[0]]
        for n in names:
            assert self._data_vars[n].dims == first.dims
        arr = Variable(('variable',) + first.dims, np.stack([self._data_vars[n].values for n in names], axis=0))
        arr._names = names
        arr._coords = self._coords
        return arr

    def transpose(self, *dims):
        new = Dataset()
        new.__class__ = self.__class__
        new._coords = OrderedDict((k, v.transpose(*dims)) for k, v in self._coords.items())
        new._data_vars = OrderedDict((k, v.transpose(*dims)) for k, v in self._data_vars.items())
        return new

    def update(self, other):
        for k, v in other._data_vars.items():
            self._data_vars[k] = v

    def mean(self, dim=None):
        new = Dataset()
        new.__class__ = self.__class__
        new._coords = OrderedDict((k, v) for k, v in self._coords.items() if dim not in v.dims)
        new._data_vars = OrderedDict((k, v.mean(dim=dim)) for k, v in self._data_vars.items())
        return new

    def _binary(self, other, op):
        new = Dataset()
        new.__class__ = self.__class__
        new._coords = self._coords
        for k, v in self._data_vars.items():
            o = other._data_vars[k] if isinstance(other, Dataset) else other
            new._data_vars[k] = v._binary(o, op)
        return new

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def to_netcdf(self, filename):
        raise NotImplementedError("netCDF is not available in this image")


def open_dataset(*a, **k):
    raise NotImplementedError("netCDF is not available in this image")
