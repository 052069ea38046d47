// miplib/cuda/mps.cpp - free-format MPS reader for the Cuda backend (bulk
// model ingress that bypasses the expression tree; the counterpart of CudaSolver::dump()).
//
// The reference has no reader of its own - its backends' readers live inside SCIP / lp_solve /
// Gurobi - so the conventions follow what those (and HiGHS, which the tests compare against) agree
// on: the first N row is the objective, further N rows are ignored; RHS on the objective row is
// minus the objective constant (dropped, like the constant of set_objective,
// miplib/lpsolve/solver.cpp:70-73); columns default to [0, +inf); RANGES widen one-sided rows
// (E rows to the side of the sign); UP with a negative value on a column whose lower bound is
// still 0 moves the lower bound to -inf; BV / LI / UI make a column integral; MARKER lines
// ('INTORG' / 'INTEND') make the columns between them integral.  Names must not contain blanks.
#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "solver.hpp"

namespace miplib {

namespace {

[[noreturn]] void fail(std::string const& what) { throw std::logic_error(what); }

[[noreturn]] void bad(std::string const& path, std::size_t line, std::string const& what)
{
  throw std::logic_error("miplib_cuda_read_mps: " + path + ":" + std::to_string(line) + ": " + what);
}

double number(std::string const& tok, std::string const& path, std::size_t line)
{
  try
  {
    std::size_t used = 0;
    double const v = std::stod(tok, &used);
    if (used != tok.size()) bad(path, line, "not a number: " + tok);
    return v;
  }
  catch (std::logic_error const&)
  {
    // 1e+30-style sentinels and "inf" spellings
    if (tok == "inf" || tok == "+inf" || tok == "Inf" || tok == "+Inf" || tok == "Infinity" || tok == "+Infinity")
      return INFINITY;
    if (tok == "-inf" || tok == "-Inf" || tok == "-Infinity") return -INFINITY;
    bad(path, line, "not a number: " + tok);
  }
}

}  // namespace

void CudaSolver::read_mps(Solver const& solver, std::string const& path)
{
  std::ifstream in(path);
  if (!in) fail("miplib_cuda_read_mps: cannot open " + path);
  double const inf = infinity();

  enum Section { None, Rows, Columns, Rhs, Ranges, Bounds, ObjSense, Done } section = None;
  struct Row { char sense; std::vector<std::pair<std::int32_t, double>> entries; double rhs = 0, range = NAN; };
  std::vector<Row> rows;                                   // constraint rows, file order
  std::unordered_map<std::string, std::int64_t> row_of;   // name -> index into rows, -1 objective, -2 ignored
  struct Col { double cost = 0, lb = 0, ub = INFINITY; bool integer = false, binary = false, bounded = false; };
  std::vector<Col> cols;
  std::unordered_map<std::string, std::int32_t> col_of;
  std::vector<std::string> col_names;
  bool maximize = false, in_int = false, have_objective = false;

  std::string text;
  std::size_t line = 0;
  while (section != Done && std::getline(in, text))
  {
    ++line;
    if (text.empty() || text[0] == '*') continue;
    std::istringstream ss(text);
    std::vector<std::string> tok;
    for (std::string t; ss >> t;) tok.push_back(t);
    if (tok.empty()) continue;
    bool const header = text[0] != ' ' && text[0] != '\t';
    if (header)
    {
      std::string const& h = tok[0];
      if (h == "NAME") continue;
      if (h == "ROWS") section = Rows;
      else if (h == "COLUMNS") section = Columns;
      else if (h == "RHS") section = Rhs;
      else if (h == "RANGES") section = Ranges;
      else if (h == "BOUNDS") section = Bounds;
      else if (h == "OBJSENSE")
      {
        section = ObjSense;
        if (tok.size() > 1) { maximize = tok[1] == "MAX" || tok[1] == "MAXIMIZE"; section = None; }
      }
      else if (h == "ENDATA") section = Done;
      else if (h == "OBJSENSEMAX") maximize = true;
      else bad(path, line, "unknown section " + h);
      continue;
    }
    switch (section)
    {
      case ObjSense:
        maximize = tok[0] == "MAX" || tok[0] == "MAXIMIZE";
        break;
      case Rows:
      {
        if (tok.size() != 2 || tok[0].size() != 1) bad(path, line, "expected '<N|L|G|E> name'");
        char const s = tok[0][0];
        if (row_of.count(tok[1])) bad(path, line, "duplicate row " + tok[1]);
        if (s == 'N')
        {
          row_of[tok[1]] = have_objective ? -2 : -1;
          have_objective = true;
        }
        else if (s == 'L' || s == 'G' || s == 'E')
        {
          row_of[tok[1]] = static_cast<std::int64_t>(rows.size());
          rows.push_back(Row{s, {}, 0.0, NAN});
        }
        else bad(path, line, "unknown row sense " + tok[0]);
        break;
      }
      case Columns:
      {
        if (tok.size() >= 3 && tok[1] == "'MARKER'")
        {
          if (tok[2] == "'INTORG'") in_int = true;
          else if (tok[2] == "'INTEND'") in_int = false;
          else bad(path, line, "unknown marker " + tok[2]);
          break;
        }
        if (tok.size() != 3 && tok.size() != 5) bad(path, line, "expected 'column row value [row value]'");
        auto it = col_of.find(tok[0]);
        if (it == col_of.end())
        {
          it = col_of.emplace(tok[0], static_cast<std::int32_t>(cols.size())).first;
          Col c;
          c.integer = in_int;
          cols.push_back(c);
          col_names.push_back(tok[0]);
        }
        for (std::size_t k = 1; k + 1 < tok.size(); k += 2)
        {
          auto r = row_of.find(tok[k]);
          if (r == row_of.end()) bad(path, line, "unknown row " + tok[k]);
          double const v = number(tok[k + 1], path, line);
          if (r->second == -1) cols[it->second].cost = v;
          else if (r->second >= 0 && v != 0) rows[r->second].entries.push_back({it->second, v});
        }
        break;
      }
      case Rhs:
      case Ranges:
      {
        // the set name is optional in free format: an even token count means it is absent
        std::size_t const first = tok.size() % 2 == 1 ? 1 : 0;
        for (std::size_t k = first; k + 1 < tok.size(); k += 2)
        {
          auto r = row_of.find(tok[k]);
          if (r == row_of.end()) bad(path, line, "unknown row " + tok[k]);
          double const v = number(tok[k + 1], path, line);
          if (r->second < 0) continue;                 // objective constant: dropped
          if (section == Rhs) rows[r->second].rhs = v;
          else rows[r->second].range = v;
        }
        break;
      }
      case Bounds:
      {
        if (tok.size() < 2 || tok.size() > 4) bad(path, line, "expected '<type> [set] column [value]'");
        std::string const& type = tok[0];
        bool const needs_value = type == "UP" || type == "LO" || type == "FX" || type == "LI" || type == "UI";
        // '<type> [set] column value' for the valued types.  BV / FR / MI / PL take no value, but some
        // writers (CPLEX, SCIP) add one anyway: '<type> [set] column [ignored]' - the column is
        // whichever of the candidate positions names one
        std::size_t name_at = tok.size() - (needs_value ? 2 : 1);
        if (!needs_value && tok.size() >= 3 && !col_of.count(tok[name_at]) && col_of.count(tok[name_at - 1]))
          name_at -= 1;                                  // trailing value field
        if (name_at < 1 || name_at > 2) bad(path, line, "malformed bound");
        auto c = col_of.find(tok[name_at]);
        if (c == col_of.end()) bad(path, line, "unknown column " + tok[name_at]);
        Col& col = cols[c->second];
        double const v = needs_value ? number(tok.back(), path, line) : 0.0;
        col.bounded = true;
        if (type == "UP" || type == "UI")
        {
          if (v < 0 && col.lb == 0) col.lb = -INFINITY;
          col.ub = v;
          if (type == "UI") col.integer = true;
        }
        else if (type == "LO" || type == "LI")
        {
          col.lb = v;
          if (type == "LI") col.integer = true;
        }
        else if (type == "FX") col.lb = col.ub = v;
        else if (type == "FR") { col.lb = -INFINITY; col.ub = INFINITY; }
        else if (type == "MI") col.lb = -INFINITY;
        else if (type == "PL") col.ub = INFINITY;
        else if (type == "BV") { col.lb = 0; col.ub = 1; col.integer = col.binary = true; }
        else bad(path, line, "unknown bound type " + type);
        break;
      }
      default:
        bad(path, line, "data line outside a section");
    }
  }
  if (section != Done) fail("miplib_cuda_read_mps: " + path + ": no ENDATA");

  // the parsed model as CSR, through the same door as miplib_cuda_load_csr
  std::size_t const first_new = m_columns.size();

  std::vector<std::int64_t> ptr{0};
  std::vector<std::int32_t> idx;
  std::vector<double> val, lo, hi, c, lb, ub;
  std::vector<std::int32_t> types;
  for (Row& r : rows)
  {
    std::sort(r.entries.begin(), r.entries.end());
    for (std::size_t k = 0; k < r.entries.size(); ++k)
    {
      if (k && r.entries[k].first == r.entries[k - 1].first) fail("miplib_cuda_read_mps: " + path + ": a coefficient is given twice");
      idx.push_back(r.entries[k].first);
      val.push_back(r.entries[k].second);
    }
    ptr.push_back(static_cast<std::int64_t>(idx.size()));
    double l = r.sense == 'L' ? -inf : r.rhs, h = r.sense == 'G' ? inf : r.rhs;
    if (!std::isnan(r.range))
    {
      double const w = std::fabs(r.range);
      if (r.sense == 'L') l = h - w;
      else if (r.sense == 'G') h = l + w;
      else if (r.range >= 0) h = l + w;
      else l = h - w;
    }
    lo.push_back(l);
    hi.push_back(h);
  }
  for (Col& col : cols)
  {
    // an integer column declared only by 'MARKER' lines, with no BOUNDS entry at all, is binary:
    // the convention of CPLEX, SCIP and HiGHS (whose reader the tests compare with)
    if (col.integer && !col.bounded) col.ub = 1;
    c.push_back(col.cost);
    lb.push_back(std::isinf(col.lb) ? (col.lb < 0 ? -inf : inf) : col.lb);
    ub.push_back(std::isinf(col.ub) ? (col.ub < 0 ? -inf : inf) : col.ub);
    bool const binary = col.integer && col.lb == 0 && col.ub == 1;
    types.push_back(binary ? 1 : col.integer ? 2 : 0);
  }
  load_csr(solver, static_cast<std::int64_t>(rows.size()), static_cast<std::int64_t>(cols.size()), ptr.data(),
           idx.data(), val.data(), c.data(), lb.data(), ub.data(), types.data(), lo.data(), hi.data(), maximize);
  for (std::size_t j = 0; j < col_names.size(); ++j) m_column_names[first_new + j] = col_names[j];
}

}  // namespace miplib
