from detectron2.modeling import Backbone  # noqa: F401
