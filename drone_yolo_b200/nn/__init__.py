from .tasks import BaseModel, DetectionModel, parse_model, yaml_model_load

__all__ = ("BaseModel", "DetectionModel", "parse_model", "yaml_model_load")
