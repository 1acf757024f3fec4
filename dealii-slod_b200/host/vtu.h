// Minimal VTU (VTK UnstructuredGrid, ASCII) writer for fields on the uniform grids of this code: the host-side
// counterpart of the reference's DataOut / write_vtu_in_parallel calls (source/LOD.cc:248-293, :1262-1377,
// include/Diffusion.h:70-108).  No deal.II: a grid of nc^dim quadrilaterals / hexahedra on the unit cube, point data on
// its (nc + 1)^dim nodes (x fastest), cell data on its cells (x fastest).  Vector fields (ncomp > 1) are written with
// three components like deal.II's component_is_part_of_vector interpretation.
#pragma once
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace vtu {

struct Field {
  std::string name;
  int ncomp;            // values per node / cell
  const double *data;   // [index * ncomp + comp]
};

inline void write_fields(std::ofstream &out, const std::vector<Field> &fields, size_t n) {
  for (const Field &f : fields) {
    const int nc_out = f.ncomp == 1 ? 1 : 3;
    out << "        <DataArray type=\"Float64\" Name=\"" << f.name << "\" NumberOfComponents=\"" << nc_out
        << "\" format=\"ascii\">\n";
    for (size_t i = 0; i < n; ++i) {
      for (int c = 0; c < nc_out; ++c) out << (c < f.ncomp ? f.data[i * f.ncomp + c] : 0.0) << ' ';
      out << '\n';
    }
    out << "        </DataArray>\n";
  }
}

inline void write(const std::string &file, int dim, int nc, const std::vector<Field> &point_data,
                  const std::vector<Field> &cell_data) {
  std::ofstream out(file);
  if (!out) throw std::runtime_error("cannot write " + file);
  out.precision(12);
  const int np = nc + 1, npz = dim == 3 ? np : 1, ncz = dim == 3 ? nc : 1;
  const size_t n_points = (size_t)np * np * npz, n_cells = (size_t)nc * nc * ncz;
  const double h = 1.0 / nc;
  out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
      << "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"" << n_points << "\" NumberOfCells=\"" << n_cells << "\">\n"
      << "      <Points>\n        <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
  for (int z = 0; z < npz; ++z)
    for (int y = 0; y < np; ++y)
      for (int x = 0; x < np; ++x) out << x * h << ' ' << y * h << ' ' << (dim == 3 ? z * h : 0.0) << '\n';
  out << "        </DataArray>\n      </Points>\n      <Cells>\n"
      << "        <DataArray type=\"Int64\" Name=\"connectivity\" format=\"ascii\">\n";
  auto node = [&](int x, int y, int z) { return ((long long)z * np + y) * np + x; };
  for (int z = 0; z < ncz; ++z)
    for (int y = 0; y < nc; ++y)
      for (int x = 0; x < nc; ++x) {
        // VTK_QUAD / VTK_HEXAHEDRON vertex order (counter-clockwise bottom face, then top face)
        out << node(x, y, z) << ' ' << node(x + 1, y, z) << ' ' << node(x + 1, y + 1, z) << ' ' << node(x, y + 1, z);
        if (dim == 3)
          out << ' ' << node(x, y, z + 1) << ' ' << node(x + 1, y, z + 1) << ' ' << node(x + 1, y + 1, z + 1) << ' '
              << node(x, y + 1, z + 1);
        out << '\n';
      }
  out << "        </DataArray>\n        <DataArray type=\"Int64\" Name=\"offsets\" format=\"ascii\">\n";
  const int vpc = dim == 3 ? 8 : 4;
  for (size_t c = 1; c <= n_cells; ++c) out << c * vpc << '\n';
  out << "        </DataArray>\n        <DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
  for (size_t c = 0; c < n_cells; ++c) out << (dim == 3 ? 12 : 9) << '\n';
  out << "        </DataArray>\n      </Cells>\n      <PointData>\n";
  write_fields(out, point_data, n_points);
  out << "      </PointData>\n      <CellData>\n";
  write_fields(out, cell_data, n_cells);
  out << "      </CellData>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n";
  if (!out) throw std::runtime_error("write error on " + file);
}

}  // namespace vtu
