"""Import-path shim: ``model.*`` of the reference, served by mivos_b200 (hot path only)."""
