#pragma once
namespace SPGrid {}
