from .PlanarFlow import PlanarFlow
from .RadialFlow import RadialFlow
from .AffineFlow import AffineFlow

FLOWS = {"planar": PlanarFlow, "radial": RadialFlow, "affine": AffineFlow}
