"""Drop-in package: the modules of the sampling hot path live here; every other module of the reference's package of
the same name (e.g. utiles.flame_utils, models.lib.base_models) is still found in the reference checkout when that is
on sys.path, so the reference's demo/ and samples/ scripts import unchanged."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
