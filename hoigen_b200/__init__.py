"""hoigen_b200 — B200-native (sm_100a) HOI scoring forward behind HOIGen's Python surface.

    from_reference(upt)      wrap an already-built reference UPT
    build_detector(...)      same signature as upt_tip_cache_model_free_finetune_distill3.build_detector (U:1712):
                             runs the reference's own build-time code, then wraps the result
"""
__version__ = "0.1.0"


def from_reference(ref_upt):
    from .detector import UPT
    return UPT.from_reference(ref_upt)


def build_detector(args, clip_cache_keys, clip_cache_values, dino_model, dino_cache_keys, dino_cache_values,
                   gen_feature_collate, gen_target_collate, gen_verb_collate, object_to_verb, class_corr,
                   object_n_verb_to_interaction, clip_model_path, num_anno, verb2interaction=None):
    """Drop-in for U:1712. Build-time work (DETR, CLIP text tower, cache construction from `args.file1`) is the
    reference's own code — it must be importable, i.e. this is called from inside the HOIGen tree."""
    try:
        from upt_tip_cache_model_free_finetune_distill3 import build_detector as _ref_build
    except ImportError as e:  # pragma: no cover
        raise ImportError("hoigen_b200.build_detector wraps the reference's build_detector; run it from the HOIGen "
                          "repository (or build the reference module yourself and call hoigen_b200.from_reference)") from e
    ref = _ref_build(args, clip_cache_keys, clip_cache_values, dino_model, dino_cache_keys, dino_cache_values,
                     gen_feature_collate, gen_target_collate, gen_verb_collate, object_to_verb, class_corr,
                     object_n_verb_to_interaction=object_n_verb_to_interaction, clip_model_path=clip_model_path,
                     num_anno=num_anno, verb2interaction=verb2interaction)
    return from_reference(ref)
