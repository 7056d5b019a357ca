"""gfa2network_b200 -- B200-native GFA -> sparse adjacency (drop-in for gfa2network's matrix path)."""
from .builders import parse_gfa
from .utils import convert_format
from .version import __version__

__all__ = ["parse_gfa", "convert_format", "__version__"]
