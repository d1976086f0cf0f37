"""Drop-in for the reference ``BLIPScoreFusion`` (src/models/uniir_blip/blip_scorefusion/blip_sf.py:11-332;
SURVEY.md 8f-3) on the B200-native kernels.

Same constructor, state-dict layout (``visual_encoder``, ``vision_proj``, ``text_encoder`` without pooler but
with its frozen cross-attention layers, ``text_proj``, their ``_m`` momentum copies, the queues, ``temp``) and
``forward(batch, alpha, encode_mbeir_batch)`` contract as the reference.  Arithmetic (restated in
oracle/blip_oracle.py::BLIPScoreFusionOracle and pinned against the real reference class): ViT class token ->
final norm -> vision_proj; BERT in mode="text" (no cross-attention) [CLS] -> text_proj; each multiplied by its
modality mask and summed (:97-172).  The momentum-distillation loss, EMA and queue update are the BLIP_FF ones
(:174-330 is the same code as blip_ff.py:118-310) and are inherited, as is the gradient-cached training step
(``engine.train_step_blip``).  The cross-attention parameters live in the flat segment that gets neither weight
decay nor gradients, so training leaves them exactly as loaded, like torch's AdamW does for frozen parameters.
"""
from __future__ import annotations

from .blip_ff import BLIPFeatureFusion


class BLIPScoreFusion(BLIPFeatureFusion):
    MODE = "sf"

    def encode_mbeir_batch(self, batch):
        id_list = batch.get("did_list") or batch.get("qid_list")
        if id_list is None:
            raise ValueError("id_list not found in batch.")
        embeddings = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch["txt_mask_batched"],
                                                  batch["image_mask_batched"], use_momentum=False)
        assert embeddings.size(0) == len(id_list), "embeddings and id_batched must have the same batch size."
        return embeddings, id_list
