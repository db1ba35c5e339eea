from segment_anything.modeling.prompt_encoder import PromptEncoder  # noqa: F401
