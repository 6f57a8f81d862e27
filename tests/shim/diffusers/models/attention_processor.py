class AttentionProcessor:  # only used as a type annotation by the reference
    pass
