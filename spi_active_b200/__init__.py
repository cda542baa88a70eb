"""spi_active_b200 — B200-native batched SysID rollout engine (drop-in `simulator=b200` backend for
LeCAR-Lab/SPI-Active's candidate-scoring hot path).  See DESIGN.md."""
__version__ = "0.1.0"
