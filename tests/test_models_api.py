"""Drop-in boundary: same factories, state_dict keys and error behaviour as the reference's src/models
(reference tests/test_encoders.py, tests/test_adaptive_mimic.py; SURVEY 8(b) key list)."""
import pytest
import torch

import two_tower_augmented_with_adaptive_mimic_mechanism_b200 as tt

GATED = {"type": "tower",
         "id_embedding": {"params": {"embedding_dim": 8, "sparse": True}},
         "feature_encoder": {"type": "mlp", "hidden_dims": [16], "activation": "relu", "output_dim": 8, "dropout": 0.15},
         "fusion": "gated", "adaptive_mimic": {"hidden_dim": 16}}


def test_state_dict_keys_match_reference_layout():
    ue = tt.build_tower_encoder(GATED, num_embeddings=10, feature_dim=5)
    ie = tt.build_tower_encoder(GATED, num_embeddings=12, feature_dim=5)
    mm = tt.AdaptiveMimicMechanism(num_users=10, num_items=12, embedding_dim=8)
    model = tt.TwoTowerModel(ue, ie, adaptive_mimic=mm)
    keys = set(model.state_dict())
    expect = set()
    for side in ("user", "item"):
        expect |= {f"{side}_encoder.embedding.weight"}
        expect |= {f"{side}_encoder.feature_encoder.network.{i}.{p}" for i in (0, 3) for p in ("weight", "bias")}
        expect |= {f"{side}_encoder.adaptive_mimic.gate_network.{i}.{p}" for i in (0, 2) for p in ("weight", "bias")}
        expect |= {f"adaptive_mimic.{side}_augmented.weight"}
    assert keys == expect
    assert ue.embedding.sparse and isinstance(ue.embedding, torch.nn.Embedding)
    assert ue.output_dim == 8 and ue.fusion == "gated"
    assert ue.adaptive_mimic.gate_network[0].weight.shape == (16, 16)


def test_no_dropout_module_shifts_the_index_like_the_reference():
    cfg = dict(GATED, feature_encoder=dict(GATED["feature_encoder"], dropout=0.0))
    enc = tt.build_tower_encoder(cfg, num_embeddings=4, feature_dim=5)
    assert "feature_encoder.network.2.weight" in enc.state_dict()


def test_linear_and_concat_variants():
    cfg = {"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}},
           "feature_encoder": {"type": "linear", "output_dim": 6}, "fusion": "concat", "output_dim": 10}
    enc = tt.build_tower_encoder(cfg, num_embeddings=4, feature_dim=5)
    assert enc.output_dim == 10 and enc.projection.weight.shape == (10, 14)
    assert "feature_encoder.network.weight" in enc.state_dict()
    assert not enc.embedding.sparse


def test_embedding_only_and_zero_feature_dim():
    enc = tt.build_tower_encoder({"type": "embedding", "params": {"embedding_dim": 4, "sparse": True}},
                                 num_embeddings=3, feature_dim=9)
    assert enc.fusion == "identity" and enc.feature_encoder is None and enc.embedding.sparse
    enc = tt.build_tower_encoder(GATED, num_embeddings=3, feature_dim=0)
    assert enc.fusion == "identity" and enc.feature_encoder is None


@pytest.mark.parametrize("cfg,exc,msg", [
    ({"type": "transformer"}, ValueError, "Unsupported encoder type"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8, "sparse": True, "max_norm": 1.0}}}, ValueError, "max_norm"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}}, "feature_encoder": {"type": "mlp", "hidden_dims": [4], "output_dim": 6}, "fusion": "gated"}, ValueError, "must equal embedding dimension"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}}, "feature_encoder": {"type": "conv"}}, ValueError, "Unsupported feature encoder type"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}}, "feature_encoder": {"type": "mlp", "activation": "swish", "hidden_dims": [4]}}, ValueError, "Unsupported activation"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}}, "feature_encoder": {"type": "identity"}}, ValueError, "Identity feature encoder"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}, "init": {"type": "orthogonal"}}}, ValueError, "Unsupported embedding init"),
    ({"type": "tower", "id_embedding": {"params": {"embedding_dim": 8}}, "fusion": "attention"}, ValueError, "Unsupported fusion strategy"),
])
def test_config_errors(cfg, exc, msg):
    with pytest.raises(exc, match=msg):
        tt.build_tower_encoder(cfg, num_embeddings=4, feature_dim=5)


def test_mimic_errors_and_deprecated_alias():
    with pytest.raises(ValueError, match="must be positive"):
        tt.AdaptiveMimicMechanism(num_users=0, num_items=3, embedding_dim=4)
    mm = tt.AdaptiveMimicMechanism(num_users=2, num_items=3, embedding_dim=4)
    with pytest.raises(ValueError, match="required for mimic"):
        mm(user_indices=None, item_indices=None, user_embedding=torch.zeros(1, 4), item_embedding=torch.zeros(1, 4))
    base = torch.zeros(1, 4)
    assert mm.augment_users(None, base) is base and mm.augment_items(None, base) is base
    with pytest.raises(ValueError, match="torch.long"):
        mm.augment_items(torch.zeros(1, dtype=torch.int32), base)
    with pytest.warns(DeprecationWarning):
        enc = tt.build_tower_encoder(dict(GATED, fusion="adaptive_mimic"), num_embeddings=4, feature_dim=5)
    assert enc.fusion == "gated"


def test_cpu_tensors_are_refused_not_silently_computed():
    enc = tt.build_tower_encoder(GATED, num_embeddings=4, feature_dim=5)
    with pytest.raises(tt.TtamError, match="no CPU path"):
        enc({"indices": torch.tensor([0, 1]), "features": torch.zeros(2, 5)})
    model = tt.TwoTowerModel(enc, enc)
    with pytest.raises(tt.TtamError, match="no CPU path"):
        tt.FusedEngine(model)
