from .unet import UNet  # noqa: F401
from .gunet import GUNet  # noqa: F401
