"""LLM relation decode engine (rows a9-a10) — placeholder until the batched OPT engine lands."""


def build_llm_engine(language_model, language_projection, device):
    raise NotImplementedError("LLM decode engine not built yet")
