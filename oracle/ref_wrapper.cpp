// pybind11 bindings for the UNMODIFIED reference operators, compiled where they
// lie under /root/reference/vinum_cpp/src (see build_ref.sh).  Mirrors the names
// exported by the reference's vinum/core/vinum_lib.cpp:20-167.  GenericHashAggregate is built
// from a copy of its header with one token patched for Arrow >= 4 (build_ref.sh).
// TEST INFRASTRUCTURE ONLY: the product (vinum_b200) never loads this module.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <arrow/python/pyarrow.h>

#include <single_numerical_hash_aggregate.h>
#include <multi_numerical_hash_aggregate.h>
#include <generic_hash_aggregate.h>
#include <one_group_aggregate.h>
#include <sort.h>
#include <table_batch_reader.h>

namespace py = pybind11;
namespace agg = vinum::operators::aggregate;
namespace srt = vinum::operators::sort;

template <typename T>
static void bind_aggregate(py::module_& m, const char* name) {
    py::class_<T>(m, name)
        .def(py::init<const std::vector<std::string>&, const std::vector<std::string>&,
                      const std::vector<agg::AggFuncDef>&>())
        .def("next", [](T& self, py::handle h) {
            auto r = arrow::py::unwrap_batch(h.ptr());
            if (!r.ok()) throw py::type_error("expected pyarrow.RecordBatch");
            self.Next(r.ValueOrDie());
        })
        .def("result", [](T& self) {
            return py::reinterpret_steal<py::object>(arrow::py::wrap_batch(self.Result()));
        });
}

PYBIND11_MODULE(ref_vinum_lib, m) {
    m.def("import_pyarrow", &arrow::py::import_pyarrow);

    py::enum_<agg::AggFuncType>(m, "AggFuncType")
        .value("COUNT_STAR", agg::AggFuncType::COUNT_STAR)
        .value("COUNT", agg::AggFuncType::COUNT)
        .value("MIN", agg::AggFuncType::MIN)
        .value("MAX", agg::AggFuncType::MAX)
        .value("SUM", agg::AggFuncType::SUM)
        .value("AVG", agg::AggFuncType::AVG)
        .export_values();
    py::enum_<srt::SortOrder>(m, "SortOrder")
        .value("ASC", srt::SortOrder::ASC)
        .value("DESC", srt::SortOrder::DESC)
        .export_values();
    py::class_<agg::AggFuncDef>(m, "AggFuncDef")
        .def(py::init<agg::AggFuncType, const std::string&, const std::string&>())
        .def_readonly("column_name", &agg::AggFuncDef::column_name)
        .def_readonly("out_col_name", &agg::AggFuncDef::out_col_name);

    bind_aggregate<agg::SingleNumericalHashAggregate>(m, "SingleNumericalHashAggregate");
    bind_aggregate<agg::MultiNumericalHashAggregate>(m, "MultiNumericalHashAggregate");
    bind_aggregate<agg::GenericHashAggregate>(m, "GenericHashAggregate");

    py::class_<agg::OneGroupAggregate>(m, "OneGroupAggregate")
        .def(py::init<const std::vector<agg::AggFuncDef>&>())
        .def("next", [](agg::OneGroupAggregate& self, py::handle h) {
            auto r = arrow::py::unwrap_batch(h.ptr());
            if (!r.ok()) throw py::type_error("expected pyarrow.RecordBatch");
            self.Next(r.ValueOrDie());
        })
        .def("result", [](agg::OneGroupAggregate& self) {
            return py::reinterpret_steal<py::object>(arrow::py::wrap_batch(self.Result()));
        });

    py::class_<srt::Sort>(m, "Sort")
        .def(py::init<const std::vector<std::string>&, const std::vector<srt::SortOrder>&>())
        .def("next", [](srt::Sort& self, py::handle h) {
            auto r = arrow::py::unwrap_batch(h.ptr());
            if (!r.ok()) throw py::type_error("expected pyarrow.RecordBatch");
            self.Next(r.ValueOrDie());
        })
        .def("sorted", [](srt::Sort& self) {
            return py::reinterpret_steal<py::object>(arrow::py::wrap_batch(self.Sorted()));
        });

    py::class_<vinum::operators::TableBatchReader>(m, "TableBatchReader")
        .def(py::init([](py::handle h) {
            auto r = arrow::py::unwrap_table(h.ptr());
            if (!r.ok()) throw py::type_error("expected pyarrow.Table");
            return new vinum::operators::TableBatchReader(r.ValueOrDie());
        }))
        .def("next", [](vinum::operators::TableBatchReader& self) -> py::object {
            auto b = self.Next();
            if (b == nullptr) return py::none();
            return py::reinterpret_steal<py::object>(arrow::py::wrap_batch(b));
        })
        .def("set_batch_size", [](vinum::operators::TableBatchReader& self, int64_t n) {
            self.SetBatchSize(n);
        });
}
