"""itn_b200 — host-side mirror of ITensorNetworks.jl's BP / simple-update API over libitn_b200.so."""
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, ITNError
from . import graphs
from .cache import (BeliefPropagationCache, Context, EdgeSequence, GateLayer, prepare_sequence, apply, apply_layer, gauge_walk, tree_gauge, tree_orthogonalize, default_bp_maxiter, default_context,
                    edge_scalars, environment, expect, expect2, inner, inner_network, loginner, logscalar, map_eigvals, message, message_diff,
                    message_residuals, norm_sqr, normalize, op, prepare_layer, rdm2, region_scalar, rescale, scalar,
                    scalar_factors_quotient, svd_batch, tebd_step, update, update_message, updated_message, vertex_scalars)
from .graphs import (NamedGraph, default_edge_sequence, edge_coloring, forest_cover, heavy_hex_eagle,
                     named_comb_tree, named_grid, named_path_graph, parallel_edge_sequence, tree_gauge_sequence)
from .network import ITensorNetwork, productstate, random_tensornetwork
from .dist import cut_edges, directed_id, gate_exchange_plan, halo_bytes_per_sweep, halo_plan, init_distributed, partition_vertices
from .partitions import PartitionMap, partition_plan, partitioned_network, tensordot
