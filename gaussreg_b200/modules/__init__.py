"""Host-side mirror of the reference's module interface for the coarse-registration forward
(geotransformer.modules.* and experiments/.../backbone.py): same class names, constructor signatures
and state_dict keys, forward-only, every tensor op a gaussreg_b200 CUDA kernel."""
from .kpconv import (KPConv, GroupNorm, UnaryBlock, LastUnaryBlock, ConvBlock, ResidualBlock, maxpool,  # noqa: F401
                     nearest_upsample, load_kernels)
from .backbone import KPConvFPN  # noqa: F401
from .transformer import (SinusoidalPositionalEmbedding, GeometricStructureEmbedding, GeometricTransformer,  # noqa: F401
                          RPEConditionalTransformer, RPETransformerLayer, TransformerLayer, AttentionOutput)
from .matching import (SuperPointMatching, LearnableLogOptimalTransport, LocalGlobalRegistration,  # noqa: F401
                       WeightedProcrustes, weighted_procrustes)
