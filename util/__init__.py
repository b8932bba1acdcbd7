"""Import-path shim: ``util.tensor_util`` of the reference, served by mivos_b200."""
