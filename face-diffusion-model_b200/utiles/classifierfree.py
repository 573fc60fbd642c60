"""utiles.classifierfree — classifier-free guidance wrapper (reference utiles/classifierfree.py:8-21)."""
from fdm_b200.modules import ClassifierFreeSampleModelBase


class ClassifierFreeSampleModel(ClassifierFreeSampleModelBase):
    def __init__(self, model, level=2.5):
        super().__init__(model, level)
