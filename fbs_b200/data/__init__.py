from .images import InpaintingMask, SRMask, ImageRestore
from .results import save_restored_images, save_chain_samples
