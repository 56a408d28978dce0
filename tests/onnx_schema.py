"""The ONNX message schema (onnx/onnx.proto, IR version 8: message names and FIELD NUMBERS as published) built as protobuf
descriptors at run time, so that Google's protobuf runtime -- an independent decoder -- can parse the files
open_duck_playground_b200/export_onnx.py writes with its own wire encoder.  TEST INFRASTRUCTURE (the `onnx` package itself is not
in this image; `google.protobuf` is).  Only the messages / fields a dense float MLP graph can contain are declared; unknown fields
would be preserved by protobuf as unknown fields and are asserted absent by the test."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto


def _msg(fd, name, fields):
    m = fd.message_type.add()
    m.name = name
    for fname, number, ftype, label, type_name in fields:
        f = m.field.add()
        f.name, f.number, f.type, f.label = fname, number, ftype, label
        if type_name:
            f.type_name = ".onnx." + type_name
    return m


def onnx_messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "onnx_subset.proto", "onnx", "proto3"
    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    _msg(fd, "TensorShapeProto_Dimension", [("dim_value", 1, F.TYPE_INT64, O, None), ("dim_param", 2, F.TYPE_STRING, O, None), ("denotation", 3, F.TYPE_STRING, O, None)])
    _msg(fd, "TensorShapeProto", [("dim", 1, F.TYPE_MESSAGE, R, "TensorShapeProto_Dimension")])
    _msg(fd, "TypeProto_Tensor", [("elem_type", 1, F.TYPE_INT32, O, None), ("shape", 2, F.TYPE_MESSAGE, O, "TensorShapeProto")])
    _msg(fd, "TypeProto", [("tensor_type", 1, F.TYPE_MESSAGE, O, "TypeProto_Tensor"), ("denotation", 6, F.TYPE_STRING, O, None)])
    _msg(fd, "ValueInfoProto", [("name", 1, F.TYPE_STRING, O, None), ("type", 2, F.TYPE_MESSAGE, O, "TypeProto"), ("doc_string", 3, F.TYPE_STRING, O, None)])
    _msg(fd, "TensorProto", [("dims", 1, F.TYPE_INT64, R, None), ("data_type", 2, F.TYPE_INT32, O, None), ("float_data", 4, F.TYPE_FLOAT, R, None),
                             ("int32_data", 5, F.TYPE_INT32, R, None), ("string_data", 6, F.TYPE_BYTES, R, None), ("int64_data", 7, F.TYPE_INT64, R, None),
                             ("name", 8, F.TYPE_STRING, O, None), ("raw_data", 9, F.TYPE_BYTES, O, None), ("double_data", 10, F.TYPE_DOUBLE, R, None),
                             ("uint64_data", 11, F.TYPE_UINT64, R, None), ("doc_string", 12, F.TYPE_STRING, O, None), ("data_location", 14, F.TYPE_INT32, O, None)])
    _msg(fd, "AttributeProto", [("name", 1, F.TYPE_STRING, O, None), ("f", 2, F.TYPE_FLOAT, O, None), ("i", 3, F.TYPE_INT64, O, None), ("s", 4, F.TYPE_BYTES, O, None),
                                ("t", 5, F.TYPE_MESSAGE, O, "TensorProto"), ("floats", 7, F.TYPE_FLOAT, R, None), ("ints", 8, F.TYPE_INT64, R, None),
                                ("strings", 9, F.TYPE_BYTES, R, None), ("doc_string", 13, F.TYPE_STRING, O, None), ("type", 20, F.TYPE_INT32, O, None),
                                ("ref_attr_name", 21, F.TYPE_STRING, O, None)])
    _msg(fd, "NodeProto", [("input", 1, F.TYPE_STRING, R, None), ("output", 2, F.TYPE_STRING, R, None), ("name", 3, F.TYPE_STRING, O, None),
                           ("op_type", 4, F.TYPE_STRING, O, None), ("attribute", 5, F.TYPE_MESSAGE, R, "AttributeProto"), ("doc_string", 6, F.TYPE_STRING, O, None),
                           ("domain", 7, F.TYPE_STRING, O, None)])
    _msg(fd, "GraphProto", [("node", 1, F.TYPE_MESSAGE, R, "NodeProto"), ("name", 2, F.TYPE_STRING, O, None), ("initializer", 5, F.TYPE_MESSAGE, R, "TensorProto"),
                            ("doc_string", 10, F.TYPE_STRING, O, None), ("input", 11, F.TYPE_MESSAGE, R, "ValueInfoProto"), ("output", 12, F.TYPE_MESSAGE, R, "ValueInfoProto"),
                            ("value_info", 13, F.TYPE_MESSAGE, R, "ValueInfoProto")])
    _msg(fd, "OperatorSetIdProto", [("domain", 1, F.TYPE_STRING, O, None), ("version", 2, F.TYPE_INT64, O, None)])
    _msg(fd, "StringStringEntryProto", [("key", 1, F.TYPE_STRING, O, None), ("value", 2, F.TYPE_STRING, O, None)])
    _msg(fd, "ModelProto", [("ir_version", 1, F.TYPE_INT64, O, None), ("producer_name", 2, F.TYPE_STRING, O, None), ("producer_version", 3, F.TYPE_STRING, O, None),
                            ("domain", 4, F.TYPE_STRING, O, None), ("model_version", 5, F.TYPE_INT64, O, None), ("doc_string", 6, F.TYPE_STRING, O, None),
                            ("graph", 7, F.TYPE_MESSAGE, O, "GraphProto"), ("opset_import", 8, F.TYPE_MESSAGE, R, "OperatorSetIdProto"),
                            ("metadata_props", 14, F.TYPE_MESSAGE, R, "StringStringEntryProto")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:                                                     # older protobuf
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return {n: get(pool.FindMessageTypeByName("onnx." + n)) for n in ("ModelProto", "GraphProto", "NodeProto", "TensorProto", "ValueInfoProto")}
