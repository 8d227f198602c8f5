"""Import shim that lets THIS container import the reference's own hot-path modules.

Only used by ``tests/golden/make_golden.py`` (run once, here, with /root/reference mounted);
never on the GPU box and never by the product.  The reference needs Python < 3.12 and the
packages commonroad-io, commonroad-drivability-checker, omegaconf, shapely, methodtools,
matplotlib ... none of which are installed.  None of them take part in the
arithmetic of ``check_feasibility`` / the polynomial classes / the active cost terms, so we
register inert stand-in modules for them and then import the *unmodified* reference files.

The three third-party functions that DO take part in the arithmetic are provided explicitly and
are the documented parity-unpinned definitions (oracle/frenet_oracle.py docstring):
``commonroad.common.util.make_valid_orientation``, the CCosy point conversion (handed in by the
caller as a coordinate-system object), and ``scipy.integrate.simps`` (removed in scipy 1.14+,
aliased to ``simpson`` whose default equals 1.13.1's).
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np

STUB_ROOTS = ("commonroad", "commonroad_dc", "commonroad_route_planner", "omegaconf", "shapely",
              "methodtools", "matplotlib", "vehiclemodels", "triangle", "onnxruntime", "pygeos",
              "imageio", "rich", "wale_net", "commonroad_rp", "PIL", "networkx", "seaborn", "pandas_stub",
              "cvxpy", "casadi", "mpl_toolkits", "prediction", "frenetix", "frenetix_occlusion",
              "wale_net_lite_stub", "toml", "tqdm_stub", "pymoo", "psutil_stub", "commonroad_helper_functions")


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(cls):
        return iter(())


class _Dummy(metaclass=_DummyMeta):
    """Anything-goes placeholder: usable as base class, callable, attribute bag."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        name = module.__name__
        if name == "methodtools":
            def lru_cache(*a, **k):
                return lambda f: f
            module.lru_cache = lru_cache
        elif name == "commonroad.common.validity":
            module.is_real_number = lambda x: isinstance(x, (int, float, np.integer, np.floating))
            module.is_natural_number = lambda x: isinstance(x, (int, np.integer)) and x >= 0
            module.is_real_number_vector = lambda x, length=None: True
            module.is_positive = lambda x: x > 0
            module.ValidTypes = types.SimpleNamespace(NUMBERS=(int, float, np.integer, np.floating))
        elif name == "commonroad.common.util":
            def make_valid_orientation(angle):
                # restated from commonroad-io 2024.2 (third-party, parity-unpinned)
                two_pi = 2.0 * np.pi
                angle = angle % two_pi
                if np.pi <= angle <= two_pi:
                    angle = angle - two_pi
                return angle
            module.make_valid_orientation = make_valid_orientation
        elif name == "omegaconf":
            class OmegaConf:
                @staticmethod
                def to_object(x):
                    return dict(x)
            module.OmegaConf = OmegaConf


def install(reference_root="/root/reference"):
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import scipy.integrate
    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = scipy.integrate.simpson
