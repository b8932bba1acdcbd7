"""Import-path shim: ``util.tensor_util`` / ``util.palette`` of the reference, served by mivos_b200."""
from pkgutil import extend_path

# A reference checkout placed AFTER this repository on sys.path keeps serving the sub-modules this package
# does not provide (e.g. interact.interaction, util.logger, model.losses): same-named modules resolve here.
__path__ = extend_path(__path__, __name__)
