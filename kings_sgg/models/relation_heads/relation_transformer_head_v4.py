"""Drop-in module path of the reference head.

``configs/psg/baseline_v4_ov.py`` registers the head through
``custom_imports = dict(imports=[..., 'kings_sgg.models.relation_heads.relation_transformer_head_v4'])``
(reference config :7-13).  Importing this module registers the B200 implementation under the same
registry name ``RelationTransformerHeadV4``, so that config runs unchanged.
"""
from openpsg_b200.categories import object_categories, relation_categories  # noqa: F401
from openpsg_b200.head import RelationTransformerHeadV4  # noqa: F401

__all__ = ["RelationTransformerHeadV4", "object_categories", "relation_categories"]
