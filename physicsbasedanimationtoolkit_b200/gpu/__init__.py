"""``pbat.gpu``: the VBD integrator (``vbd``) and the pieces of its contact path that the reference also exposes on their
own (``geometry.Aabb`` / ``geometry.Bvh``, ``contact.VertexTriangleMixedCcdDcd``, ``common.Buffer``)."""
from . import common, contact, geometry, vbd, xpbd  # noqa: F401
