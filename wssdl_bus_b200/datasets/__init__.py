"""datasets/ twin: only the evaluation that consumes the gathered detections
(datasets/voc_eval_bus.py); image databases, XML parsing and result files are out of scope."""
