"""EDF ingest for the GPU path: recordings arrive as int16
records; shipping them over PCIe as int16 and applying the per-channel
calibration on the device moves a quarter of the bytes a float64 chunk does."""
