"""B200-native AttFind hot path of StylEx (drop-in for the reference's Python surface).

See DESIGN.md.  Host-side mirror of the reference interface: ``modules`` (Generator, GeneratorBlock,
RGBBlock, Conv2DMod, Blur), ``attfind`` (attfind_extraction, find_significant_styles, ...), ``counterfactual``
(generate_change_image_given_dlatent, generate_images_given_dlatent, visualize_style), ``classifiers``
(ResNet / MobileNet wrappers, PyTorch), ``training`` (the loss helpers and the step of ``Trainer.train``), ``stylex`` (StylEx container, encoder / discriminator, the reference's
checkpoint format, batched phase A), ``dist`` (latent sharding + the one all-gather).  The CUDA kernels and
the C-ABI library live under ``csrc`` (declared in ``include/stylex_b200.h``, bound in ``_native``).
"""
from . import synthetic  # noqa: F401
from . import _native  # noqa: F401
from . import training  # noqa: F401
from .modules import (Blur, Conv2DMod, Conv2DModFunction, Generator, GeneratorBlock, GeneratorPlan, RGBBlock, image_noise,  # noqa: F401
                      styles_def_to_tensor)
from .classifiers import MobileNet, ResNet, make_classifier  # noqa: F401
from .attfind import (attfind_extraction, attfind_select, attfind_sweep, attfind_verify_topk, filter_unstable_images,  # noqa: F401
                      find_significant_styles, get_min_max_style_vectors, load_records, save_records,
                      sindex_to_block_idx_and_index)
from .counterfactual import (draw_on_image, generate_change_image_given_dlatent, generate_images_given_dlatent,  # noqa: F401
                             render_counterfactuals, visualize_style, visualize_style_by_distance_in_s)
from .stylex import (DiscriminatorBlock, DiscriminatorE, EqualLinear, StyleVectorizer, StylEx, encode_images,  # noqa: F401
                     find_discriminator_threshold, load_checkpoint, load_stylex, model_loader, save_checkpoint,
                     stylex_config)

__version__ = "0.2.0"
