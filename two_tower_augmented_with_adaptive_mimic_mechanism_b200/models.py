"""Drop-in mirror of the reference's `src/models` package (encoders.py, adaptive_mimic.py, two_tower.py).

Same factory / class names, constructor arguments, error behaviour and — because checkpoints and
`_collect_parameter_groups` (reference training.py:276-309) address parameters by name —
the same `state_dict()` keys:

    {user,item}_encoder.embedding.weight
    {user,item}_encoder.feature_encoder.network[.{i}].{weight,bias}
    {user,item}_encoder.adaptive_mimic.gate_network.{0,2}.{weight,bias}
    {user,item}_encoder.projection.{weight,bias}
    adaptive_mimic.{user,item}_augmented.weight

The parameter containers are ordinary torch modules (that is what fixes the key names); the arithmetic
is not: every forward/backward goes through the sm_100a kernels of libttam.so (`tower_ops`), on CUDA
tensors only.  There is no CPU path — a CPU tensor raises `TtamError`.
"""
from __future__ import annotations

import warnings
from typing import Any, Mapping, Optional

import torch
from torch import nn

from . import tower_ops

_ACTIVATIONS = {"relu": nn.ReLU, "gelu": nn.GELU, "tanh": nn.Tanh, "selu": nn.SELU}
_FUSIONS = ("identity", "sum", "concat", "gated")


# ------------------------------------------------------------------------------------------------
# ID embedding (reference encoders.py:19-65)
# ------------------------------------------------------------------------------------------------
def _initialise_table(table: nn.Embedding, spec: Optional[Mapping[str, Any]]) -> None:
    spec = spec or {"type": "normal", "std": 0.02}
    kind = str(spec.get("type", "normal")).lower()
    w = table.weight
    if kind == "normal":
        nn.init.normal_(w, mean=0.0, std=float(spec.get("std", 0.02)))
    elif kind == "uniform":
        b = float(spec.get("bound", 0.1))
        nn.init.uniform_(w, -b, b)
    elif kind == "xavier_normal":
        nn.init.xavier_normal_(w)
    elif kind == "xavier_uniform":
        nn.init.xavier_uniform_(w)
    else:
        raise ValueError(f"Unsupported embedding init type: {kind}")


def build_id_embedding(config: Mapping[str, Any], *, num_embeddings: int,
                       device: torch.device | None = None) -> nn.Embedding:
    """reference encoders.py:39-65.  The table stays an `nn.Embedding` because the training loop
    checks `isinstance(encoder.embedding, nn.Embedding)` and reads `.sparse` (training.py:291-297)."""
    params = config.get("params", {})
    sparse = bool(params.get("sparse", False))
    max_norm = params.get("max_norm")
    if sparse and max_norm is not None:
        raise ValueError("max_norm is not supported when using sparse embeddings.")
    table = nn.Embedding(num_embeddings=num_embeddings, embedding_dim=int(params.get("embedding_dim", 64)),
                         padding_idx=params.get("padding_idx"), max_norm=max_norm, sparse=sparse)
    _initialise_table(table, config.get("init"))
    return table if device is None else table.to(device)


# ------------------------------------------------------------------------------------------------
# feature encoder (reference encoders.py:68-146)
# ------------------------------------------------------------------------------------------------
def _get_activation(name: str) -> nn.Module:
    try:
        return _ACTIVATIONS[name.lower()]()
    except KeyError:
        raise ValueError(f"Unsupported activation '{name}'") from None


class FeatureEncoderWrapper(nn.Module):
    """Parameter container of the metadata projection; `network` keeps the reference's key layout
    (`network.weight` for linear, `network.{0,3,..}` for the MLP).  `kind`, `activation`, `dropout` are
    what the fused kernels need to know."""

    def __init__(self, network: nn.Module, output_dim: int, *, kind: str = "custom", activation: str = "relu",
                 dropout: float = 0.0) -> None:
        super().__init__()
        self.network = network
        self.output_dim = output_dim
        self.kind = kind
        self.activation = activation
        self.dropout = float(dropout)

    def linear_layers(self) -> list[nn.Linear]:
        if isinstance(self.network, nn.Linear):
            return [self.network]
        if isinstance(self.network, nn.Sequential):
            return [m for m in self.network if isinstance(m, nn.Linear)]
        return []

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return tower_ops.feature_encoder_forward(self, inputs)


def build_feature_encoder(config: Mapping[str, Any] | None, *, input_dim: int,
                          fallback_output_dim: int) -> FeatureEncoderWrapper | None:
    if input_dim == 0:
        return None
    cfg = dict(config or {})
    unknown = set(cfg) - {"type", "output_dim", "hidden_dims", "activation", "dropout"}
    if unknown:
        raise TypeError(f"unexpected feature encoder option(s): {sorted(unknown)}")
    kind = cfg.get("type", "linear")
    out_dim = int(cfg.get("output_dim") or fallback_output_dim)
    act = cfg.get("activation", "relu")
    p = cfg.get("dropout", 0.0) or 0.0
    if kind == "identity":
        if input_dim != out_dim:
            raise ValueError("Identity feature encoder requires input_dim == output_dim.")
        return FeatureEncoderWrapper(nn.Identity(), out_dim, kind="identity")
    if kind == "linear":
        lin = nn.Linear(input_dim, out_dim)
        nn.init.xavier_uniform_(lin.weight)
        return FeatureEncoderWrapper(lin, out_dim, kind="linear")
    if kind == "mlp":
        act_module = _get_activation(act)
        stack: list[nn.Module] = []
        width = input_dim
        for h in [int(h) for h in (cfg.get("hidden_dims") or [])]:
            lin = nn.Linear(width, h)
            nn.init.xavier_uniform_(lin.weight)
            stack += [lin, act_module]
            if p:
                stack.append(nn.Dropout(p=p))
            width = h
        head = nn.Linear(width, out_dim)
        nn.init.xavier_uniform_(head.weight)
        stack.append(head)
        return FeatureEncoderWrapper(nn.Sequential(*stack), out_dim, kind="mlp", activation=str(act).lower(), dropout=p)
    raise ValueError(f"Unsupported feature encoder type: {kind}")


# ------------------------------------------------------------------------------------------------
# gate + tower (reference encoders.py:149-331)
# ------------------------------------------------------------------------------------------------
class FeatureFusionGate(nn.Module):
    """g = sigmoid(G2 relu(G1 [e;f] + c1) + c2);  out = g*e + (1-g)*f   (reference encoders.py:149-168)."""

    def __init__(self, dim: int, hidden_dim: int | None = None) -> None:
        super().__init__()
        width = hidden_dim or dim
        self.dim = dim
        self.gate_network = nn.Sequential(nn.Linear(2 * dim, width), nn.ReLU(), nn.Linear(width, dim), nn.Sigmoid())

    def forward(self, id_repr: torch.Tensor, feature_repr: torch.Tensor) -> torch.Tensor:
        return tower_ops.gate_forward(self, id_repr, feature_repr)


class TowerEncoder(nn.Module):
    """ID embedding (+ metadata encoder, + fusion).  `forward({"indices", "features"?}) -> [R, D]`."""

    def __init__(self, *, embedding: nn.Embedding, feature_encoder: FeatureEncoderWrapper | None, fusion: str,
                 output_dim: int | None, adaptive_mimic: FeatureFusionGate | None) -> None:
        super().__init__()
        self.embedding = embedding
        self.feature_encoder = feature_encoder
        self.adaptive_mimic = adaptive_mimic
        self.num_embeddings = embedding.num_embeddings
        self.id_dim = embedding.embedding_dim
        mode = fusion
        if mode == "adaptive_mimic":
            warnings.warn("TowerEncoder fusion='adaptive_mimic' is deprecated; use fusion='gated' instead.",
                          DeprecationWarning, stacklevel=2)
            mode = "gated"
        if mode not in _FUSIONS:
            raise ValueError(f"Unsupported fusion strategy: {fusion}")
        self.fusion = "identity" if feature_encoder is None else mode
        self.output_dim = self.id_dim
        if self.fusion == "concat":
            width = self.id_dim + feature_encoder.output_dim
            self.output_dim = int(output_dim or width)
            self.projection = nn.Linear(width, self.output_dim)
            nn.init.xavier_uniform_(self.projection.weight)

    def forward(self, inputs: Mapping[str, torch.Tensor]) -> torch.Tensor:
        return tower_ops.tower_module_forward(self, inputs["indices"], inputs.get("features"))


def build_tower_encoder(config: Mapping[str, Any] | None, *, num_embeddings: int, feature_dim: int,
                        device: torch.device | None = None) -> TowerEncoder:
    cfg = config or {}
    kind = str(cfg.get("type", "tower")).lower()
    if kind not in ("tower", "embedding"):
        raise ValueError(f"Unsupported encoder type: {kind}")
    if kind == "embedding":
        table = build_id_embedding({"params": cfg.get("params", {}), "init": cfg.get("init")},
                                   num_embeddings=num_embeddings, device=device)
        return TowerEncoder(embedding=table, feature_encoder=None, fusion="identity", output_dim=None,
                            adaptive_mimic=None).to(device)
    id_cfg = cfg.get("id_embedding", {})
    table = build_id_embedding({"params": id_cfg.get("params", {}), "init": id_cfg.get("init")},
                               num_embeddings=num_embeddings, device=device)
    fusion = str(cfg.get("fusion", "gated" if feature_dim > 0 else "identity")).lower()
    encoder = build_feature_encoder(cfg.get("feature_encoder"), input_dim=feature_dim,
                                    fallback_output_dim=table.embedding_dim)
    needs_match = fusion in ("sum", "adaptive_mimic", "gated")
    if needs_match and encoder is not None and encoder.output_dim != table.embedding_dim:
        raise ValueError("Feature encoder output dimension must equal embedding dimension for 'sum' or 'gated' fusion.")
    gate = None
    if fusion in ("adaptive_mimic", "gated"):
        gate = FeatureFusionGate(dim=table.embedding_dim, hidden_dim=cfg.get("adaptive_mimic", {}).get("hidden_dim"))
    tower = TowerEncoder(embedding=table, feature_encoder=encoder, fusion=fusion, output_dim=cfg.get("output_dim"),
                         adaptive_mimic=gate)
    return tower if device is None else tower.to(device)


# ------------------------------------------------------------------------------------------------
# adaptive mimic mechanism (reference adaptive_mimic.py:20-105)
# ------------------------------------------------------------------------------------------------
class AdaptiveMimicMechanism(nn.Module):
    """Per-user / per-item augmentation tables; o = t + A[idx]; mimic losses on positive pairs."""

    def __init__(self, *, num_users: int, num_items: int, embedding_dim: int, init_std: float = 0.02) -> None:
        super().__init__()
        if num_users <= 0 or num_items <= 0:
            raise ValueError("num_users and num_items must be positive.")
        self.embedding_dim = int(embedding_dim)
        self.user_augmented = nn.Embedding(num_users, self.embedding_dim)
        self.item_augmented = nn.Embedding(num_items, self.embedding_dim)
        for table in (self.user_augmented, self.item_augmented):
            nn.init.normal_(table.weight, mean=0.0, std=init_std)

    def forward(self, *, user_indices, item_indices, user_embedding, item_embedding):
        if user_indices is None or item_indices is None:
            raise ValueError("user_indices and item_indices are required for mimic.")
        return tower_ops.mimic_forward(self, user_indices, item_indices, user_embedding, item_embedding)

    def augment_users(self, indices: Optional[torch.Tensor], base_embedding: torch.Tensor) -> torch.Tensor:
        if indices is None:
            return base_embedding
        return tower_ops.augment(self.user_augmented, indices, base_embedding)[0]

    def augment_items(self, indices: Optional[torch.Tensor], base_embedding: torch.Tensor) -> torch.Tensor:
        if indices is None:
            return base_embedding
        return tower_ops.augment(self.item_augmented, indices, base_embedding)[0]


# ------------------------------------------------------------------------------------------------
# container (reference two_tower.py:19-95)
# ------------------------------------------------------------------------------------------------
def _indices_of(inputs: Any) -> torch.Tensor | None:
    if isinstance(inputs, torch.Tensor):
        return inputs
    if isinstance(inputs, dict) and isinstance(inputs.get("indices"), torch.Tensor):
        return inputs["indices"]
    return None


class TwoTowerModel(nn.Module):
    def __init__(self, user_encoder: nn.Module, item_encoder: nn.Module, similarity: nn.Module | None = None,
                 adaptive_mimic: AdaptiveMimicMechanism | None = None) -> None:
        super().__init__()
        self.user_encoder = user_encoder
        self.item_encoder = item_encoder
        self.similarity = similarity or nn.CosineSimilarity(dim=-1)
        self.adaptive_mimic = adaptive_mimic

    def forward(self, user_inputs: Any, item_inputs: Any, *, return_embeddings: bool = False) -> dict[str, torch.Tensor]:
        u = self.user_encoder(user_inputs)
        i = self.item_encoder(item_inputs)
        out: dict[str, torch.Tensor] = {}
        loss_u = loss_i = None
        if self.adaptive_mimic is not None:
            u, i, loss_u, loss_i = self.adaptive_mimic(user_indices=_indices_of(user_inputs),
                                                       item_indices=_indices_of(item_inputs),
                                                       user_embedding=u, item_embedding=i)
        if return_embeddings:
            out["user_embedding"], out["item_embedding"] = u, i
        if loss_u is not None:
            out["mimic_user_loss"] = loss_u
        if loss_i is not None:
            out["mimic_item_loss"] = loss_i
        out["score"] = self.similarity(u, i)
        return out
