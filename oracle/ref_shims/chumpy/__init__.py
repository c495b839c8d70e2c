"""Pickle stand-in for chumpy==0.70 (absent): ``flame.pkl`` stores ``shapedirs`` etc. as ``chumpy.ch.Ch`` objects; the
reference's ``get_flame_model`` (model/utils.py:84-89) unpickles them and ``smplx.utils.to_np`` turns them into arrays via
``np.array(obj)``.  ``Ch`` here keeps the pickled state and exposes the payload ``x`` through ``__array__``."""
from .ch import Ch  # noqa: F401
