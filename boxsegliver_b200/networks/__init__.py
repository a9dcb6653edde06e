from .unet import UNet  # noqa: F401
from .gunet import GUNet  # noqa: F401
from .unet3d import UNet3D  # noqa: F401
from .unetinter import UNetInter  # noqa: F401
