"""Object / relation vocabularies of the PSG dataset (data, not code).

The reference head imports these two lists from
``kings_sgg/models/detectors/mask2former_relation_v2.py:33-37``; they are the 133 COCO-panoptic
class names (80 things then 53 stuff, with the ``-stuff`` / ``-merged`` / ``-other`` suffixes
stripped, ``mask2former_relation_v2.py:22-30``) and the 56 PSG predicates.
"""

_COCO_PANOPTIC_RAW = (
    # 80 thing classes
    "person|bicycle|car|motorcycle|airplane|bus|train|truck|boat|traffic light|fire hydrant|stop sign|"
    "parking meter|bench|bird|cat|dog|horse|sheep|cow|elephant|bear|zebra|giraffe|backpack|umbrella|"
    "handbag|tie|suitcase|frisbee|skis|snowboard|sports ball|kite|baseball bat|baseball glove|skateboard|"
    "surfboard|tennis racket|bottle|wine glass|cup|fork|knife|spoon|bowl|banana|apple|sandwich|orange|"
    "broccoli|carrot|hot dog|pizza|donut|cake|chair|couch|potted plant|bed|dining table|toilet|tv|laptop|"
    "mouse|remote|keyboard|cell phone|microwave|oven|toaster|sink|refrigerator|book|clock|vase|scissors|"
    "teddy bear|hair drier|toothbrush|"
    # 53 stuff classes
    "banner|blanket|bridge|cardboard|counter|curtain|door-stuff|floor-wood|flower|fruit|gravel|house|light|"
    "mirror-stuff|net|pillow|platform|playingfield|railroad|river|road|roof|sand|sea|shelf|snow|stairs|tent|"
    "towel|wall-brick|wall-stone|wall-tile|wall-wood|water-other|window-blind|window-other|tree-merged|"
    "fence-merged|ceiling-merged|sky-other-merged|cabinet-merged|table-merged|floor-other-merged|"
    "pavement-merged|mountain-merged|grass-merged|dirt-merged|paper-merged|food-other-merged|"
    "building-other-merged|rock-merged|wall-other-merged|rug-merged"
)


def _clean(name: str) -> str:
    for suffix in ("-stuff", "-merged", "-other"):
        name = name.replace(suffix, "")
    return name


object_categories = [_clean(n) for n in _COCO_PANOPTIC_RAW.split("|")]

relation_categories = (
    "over|in front of|beside|on|in|attached to|hanging from|on back of|falling off|going down|painted on|"
    "walking on|running on|crossing|standing on|lying on|sitting on|flying over|jumping over|jumping from|"
    "wearing|holding|carrying|looking at|guiding|kissing|eating|drinking|feeding|biting|catching|picking|"
    "playing with|chasing|climbing|cleaning|playing|touching|pushing|pulling|opening|cooking|talking to|"
    "throwing|slicing|driving|riding|parked on|driving on|about to hit|kicking|swinging|entering|exiting|"
    "enclosing|leaning on"
).split("|")

INSTANCE_OFFSET = 1000  # mmdet.core.INSTANCE_OFFSET (relation_transformer_head_v4.py:12,138)

assert len(object_categories) == 133 and len(relation_categories) == 56
