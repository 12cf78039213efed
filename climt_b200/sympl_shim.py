"""Minimal stand-in for the slice of sympl the radiation components use.

sympl is a third-party dependency of climt (`sympl>=0.5.0`, setup.py:49) and is not installable in
this image.  When the real package is importable it is used unchanged; otherwise this shim provides
just enough of `TendencyComponent.__call__` (dimension flattening to ["mid_levels", "*"], a small unit
table, re-wrapping of outputs) for the drop-in components to be exercised through `component(state)`.
"""
import numpy as np

try:  # pragma: no cover - not available in the build image
    from sympl import TendencyComponent, ImplicitTendencyComponent, DiagnosticComponent, Stepper, DataArray, initialize_numpy_arrays_with_properties  # noqa: F401
    HAVE_SYMPL = True
except Exception:  # ImportError or a broken install
    HAVE_SYMPL = False

    class DataArray:
        """Just enough of xarray.DataArray: values, dims, attrs['units']."""

        def __init__(self, values, dims=(), attrs=None):
            self.values = np.asarray(values)
            self.dims = tuple(dims)
            self.attrs = dict(attrs or {})

        @property
        def shape(self):
            return self.values.shape

        def to_units(self, units):
            return DataArray(self.values * _unit_factor(self.attrs.get("units", ""), units), self.dims,
                             {**self.attrs, "units": units})

    _CANON = {"mbar": ("Pa", 100.0), "hPa": ("Pa", 100.0), "Pa": ("Pa", 1.0),
              "g m^-2": ("kg m^-2", 1e-3), "kg m^-2": ("kg m^-2", 1.0),
              "micrometer": ("m", 1e-6), "m": ("m", 1.0),
              "degK": ("K", 1.0), "K": ("K", 1.0),
              "g/g": ("1", 1.0), "kg/kg": ("1", 1.0), "m s^-1": ("m s^-1", 1.0), "kg m^-2 s^-1": ("kg m^-2 s^-1", 1.0), "dimensionless": ("1", 1.0), "mole/mole": ("1", 1.0),
              "": ("1", 1.0), "1": ("1", 1.0),
              "W m^-2": ("W m^-2", 1.0), "degK day^-1": ("K day^-1", 1.0), "K day^-1": ("K day^-1", 1.0),
              "radians": ("rad", 1.0), "degrees": ("rad", np.pi / 180.0),
              "degrees_north": ("degrees_north", 1.0), "degrees_east": ("degrees_east", 1.0),
              "J kg^-1 degK^-1": ("J kg^-1 K^-1", 1.0), "J kg^-1 K^-1": ("J kg^-1 K^-1", 1.0), "kg m^-3": ("kg m^-3", 1.0),
              "degK s^-1": ("K s^-1", 1.0), "K s^-1": ("K s^-1", 1.0)}

    def _unit_factor(src, dst):
        if src == dst:
            return 1.0
        try:
            (b0, f0), (b1, f1) = _CANON[src], _CANON[dst]
        except KeyError as e:
            raise ValueError(f"sympl shim: unknown unit {e}") from None
        if b0 != b1:
            raise ValueError(f"sympl shim: cannot convert {src!r} to {dst!r}")
        return f0 / f1

    def _to_raw(da, dims, units):
        """DataArray -> numpy in the component's dims; '*' collects every other dim (C order)."""
        numeric = da.values.dtype.kind in "fiub"  # string-valued quantities (area_type) pass through unconverted
        vals = da.values * _unit_factor(da.attrs.get("units", ""), units) if numeric else da.values
        named = [d for d in dims if d != "*"]
        for d in named:
            if d not in da.dims:
                raise ValueError(f"dimension {d} missing from {da.dims}")
        star = [d for d in da.dims if d not in named]
        order = []
        for d in dims:
            order += [da.dims.index(x) for x in star] if d == "*" else [da.dims.index(d)]
        v = np.transpose(vals, order)
        shape, k = [], 0
        for d in dims:
            if d == "*":
                n = int(np.prod([da.values.shape[da.dims.index(x)] for x in star])) if star else 1
                shape.append(n)
                k += len(star)
            else:
                shape.append(v.shape[k])
                k += 1
        star_shape = tuple(da.values.shape[da.dims.index(x)] for x in star)
        return np.ascontiguousarray(v.reshape(shape), dtype=np.float64 if numeric else None), tuple(star), star_shape

    class TendencyComponent:
        input_properties = {}
        tendency_properties = {}
        diagnostic_properties = {}
        _diagnostic_only = False
        _stepper = False

        def __init__(self, **kwargs):
            if kwargs:
                raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")

        def __call__(self, state, *extra):
            raw, star, star_shape = {}, (), ()
            # sympl hands array_call the raw arrays under a quantity's alias when its properties declare one, and accepts outputs
            # under the alias of ANY property dict of the component (GrayLongwaveRadiation returns its tendency as "sl", the alias
            # of the input air_temperature, climt/_components/radiation.py:27-62,108)
            alias_of = {}
            for props in (self.input_properties, self.tendency_properties, self.diagnostic_properties,
                          getattr(self, "output_properties", {})):
                for name, prop in props.items():
                    if "alias" in prop:
                        alias_of[prop["alias"]] = name
            for name, prop in self.input_properties.items():
                if name not in state:
                    raise KeyError(f"state is missing input quantity {name!r}")
                raw[prop.get("alias", name)], s, ss = _to_raw(state[name], prop["dims"], prop.get("units", ""))
                if "*" in prop["dims"] and len(s) >= len(star):
                    star, star_shape = s, ss
            if "time" in state:
                raw["time"] = state["time"]
            result = self.array_call(raw, *extra)
            if self._diagnostic_only:
                tend, diag = {}, result
            elif self._stepper:
                diag, tend = result
            else:
                tend, diag = result

            def wrap(arr, prop, name=None):
                dims, shape, k = [], [], 0
                # sympl: a tendency without explicit dims takes the dims of the input quantity of the same name
                pdims = prop["dims"] if "dims" in prop else self.input_properties[name]["dims"]
                for d in pdims:
                    if d == "*":
                        dims += list(star)
                        shape += list(star_shape)
                    else:
                        dims.append(d)
                        shape.append(arr.shape[k])
                    k += 1
                return DataArray(np.asarray(arr).reshape(shape), dims, {"units": prop.get("units", "")})
            diag = {alias_of.get(k, k): v for k, v in diag.items()}
            tend = {alias_of.get(k, k): v for k, v in tend.items()}
            diag = {k: wrap(v, self.diagnostic_properties[k]) for k, v in diag.items()}
            if self._diagnostic_only:
                return diag
            return {k: wrap(v, self.tendency_properties[k], k) for k, v in tend.items()}, diag

    class DiagnosticComponent(TendencyComponent):
        """sympl.DiagnosticComponent: `component(state)` -> diagnostics dict; array_call returns the diagnostics only."""
        _diagnostic_only = True

    class Stepper(TendencyComponent):
        """sympl.Stepper: `component(state, timestep)` -> (diagnostics, new_state); array_call returns them in that order and
        `output_properties` quantities take the dims of the input of the same name."""
        output_properties = {}
        _stepper = True

        def __call__(self, state, timestep):
            saved = self.tendency_properties
            self.tendency_properties = self.output_properties
            try:
                new_state, diag = TendencyComponent.__call__(self, state, timestep)
            finally:
                self.tendency_properties = saved
            return diag, new_state

    class ImplicitTendencyComponent(TendencyComponent):
        """sympl.ImplicitTendencyComponent: `component(state, timestep)` -> array_call(raw_state, timestep)."""

        def __call__(self, state, timestep):
            return TendencyComponent.__call__(self, state, timestep)

    def initialize_numpy_arrays_with_properties(output_properties, raw_input_state, input_properties):
        dim_len = {}
        for name, prop in input_properties.items():
            if name in raw_input_state:
                for d, n in zip(prop["dims"], np.shape(raw_input_state[name])):
                    dim_len[d] = n
        out = {}
        for name, prop in output_properties.items():
            dims = prop["dims"] if "dims" in prop else input_properties[name]["dims"]   # sympl: dims of the input of the same name
            out[name] = np.zeros([dim_len[d] for d in dims], dtype=np.float64)
        return out
