"""Drop-in `llmrankers` package for the Flan-T5 hot path of ielab/llm-rankers, backed by the sm_100a engine
(libb200rank.so) instead of transformers' T5ForConditionalGeneration. Same class names, constructor keywords,
rerank()/compare()/truncate() behaviour and counters as the reference package of the same name."""
