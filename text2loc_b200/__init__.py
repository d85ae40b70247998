"""text2loc_b200 -- B200-native coarse cell-retrieval engine for Text2Loc's global
place-recognition path (drop-in for CellRetrievalNetwork.encode_text / encode_objects and
eval_epoch / run_coarse; see DESIGN.md and INTEGRATION.md).

Importing the package is cheap and works without a GPU (synthetic data, weight folding, data
packing).  Anything that computes goes through the CUDA extension and raises if it is missing.
"""

__all__ = ["CellRetrievalNetwork", "CrossMatch", "Engine", "EngineError", "eval_epoch", "run_coarse", "run_fine"]


def __getattr__(name):
    if name == "CellRetrievalNetwork":
        from .cell_retrieval import CellRetrievalNetwork

        return CellRetrievalNetwork
    if name == "CrossMatch":
        from .cross_matcher import CrossMatch

        return CrossMatch
    if name in ("Engine", "EngineError"):
        from . import engine

        return getattr(engine, name)
    if name in ("eval_epoch", "run_coarse", "run_fine"):
        from . import evaluation

        return getattr(evaluation, name)
    raise AttributeError(name)
