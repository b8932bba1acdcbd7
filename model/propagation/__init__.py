"""Import-path shim: ``model.propagation.*`` of the reference, served by mivos_b200 (hot path only)."""
from pkgutil import extend_path

# A reference checkout placed AFTER this repository on sys.path keeps serving the sub-modules this package
# does not provide (e.g. interact.interaction, util.logger, model.losses): same-named modules resolve here.
__path__ = extend_path(__path__, __name__)
