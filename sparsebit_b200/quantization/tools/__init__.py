from .calibration import CalibrationRunner, trace_quant_model  # noqa: F401
