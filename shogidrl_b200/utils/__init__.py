from .policy_mapper import PolicyOutputMapper, index_to_move, move_to_index  # noqa: F401

__all__ = ["PolicyOutputMapper", "index_to_move", "move_to_index"]
