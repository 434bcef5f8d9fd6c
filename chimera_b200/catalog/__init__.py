from .catalog import empty_catalog, pixelated_catalog
from .completeness import dVdz_completeness
from . import completeness
